//! rustlight_b200.rs -- the binding a rustlight maintainer adds as `src/b200.rs` (+ `pub mod b200;` in lib.rs, feature "b200").
//! UNCOMPILED here: this image has no Rust toolchain.  tests/test_shim_matches_header.py checks every
//! `#[repr(C)]` struct and every `extern "C"` declaration below against include/rl_b200.h, field by field.
//!
//! It declares the whole C ABI of include/rl_b200.h and implements `Integrator` for `path`, `direct` and `ao` by flattening
//! `Scene` into `rl_scene_desc`, calling `rl_render` and wrapping the result into a `BufferCollection` -- replacing the one-line
//! `compute_mc(self, sampler, accel, scene)` bodies of src/integrators/explicit/path.rs:187-196, src/integrators/direct.rs:10-19
//! and src/integrators/ao.rs.  The `accel` argument is ignored (the library builds its own structures).
//!
//! The context and the device scene stay alive across `compute()` calls (keyed on the `Scene` address), so the averaging
//! wrappers (`-a`, `-e`: src/integrators/avg.rs:45-65, equal_time.rs:21-47) re-render a resident scene, and every call advances
//! the sample offset by `nb_samples`, so successive passes draw fresh samples -- what the reference gets from the master
//! sampler's state moving on in `generate_img_blocks` (src/integrators/mod.rs:351-374).
//!
//! Everything the reference does not expose yet is in `mod patch` at the end of this file, as the (small) additions a maintainer
//! applies: `BSDF::describe`, `Emitter::describe`, two `Camera` accessors and `IndependentSampler::seed`.
#![allow(non_camel_case_types, dead_code)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};
use std::sync::Mutex;

#[repr(C)] pub struct rl_ctx { _p: [u8; 0] }
#[repr(C)] pub struct rl_scene { _p: [u8; 0] }

pub const RL_B200_ABI_VERSION: c_int = 6;

#[repr(C)] #[derive(Clone, Copy)]
pub struct rl_texture {
    pub kind: u32, pub width: u32, pub height: u32, pub pixels: *const f32, pub color0: [f32; 3], pub color1: [f32; 3],
    pub line_width: f32, pub offset: [f32; 2], pub scale: [f32; 2],
}
#[repr(C)] #[derive(Clone, Copy)]
pub struct rl_material {
    pub kind: u32, pub kd: [f32; 3], pub ks: [f32; 3], pub exponent: f32, pub weight_specular: f32,
    pub kt: [f32; 3], pub eta: [f32; 3], pub k: [f32; 3], pub ior: f32, pub alpha: f32, pub microfacet: u32, pub kd_texture: u32,
    pub ks_texture: u32, pub kt_texture: u32, pub eta_texture: u32, pub k_texture: u32,
    pub blend_a: u32, pub blend_b: u32, pub blend_weight: f32,
}
#[repr(C)]
pub struct rl_mesh_desc {
    pub p: *const f32, pub nverts: u32, pub idx: *const u32, pub ntris: u32,
    pub n: *const f32, pub uv: *const f32, pub mat: rl_material, pub emission_kind: u32, pub emission: [f32; 3], pub emission_texture: u32,
}
#[repr(C)] pub struct rl_camera_desc { pub width: u32, pub height: u32, pub sample_to_camera: [f32; 16], pub to_world: [f32; 16] }
#[repr(C)] #[derive(Clone, Copy)] pub struct rl_light_desc { pub kind: u32, pub intensity: [f32; 3], pub v: [f32; 3] }
#[repr(C)] pub struct rl_scene_desc {
    pub nmeshes: u32, pub meshes: *const rl_mesh_desc, pub camera: rl_camera_desc, pub has_volume: u32, pub has_environment: u32,
    pub nlights: u32, pub lights: *const rl_light_desc, pub ntextures: u32, pub textures: *const rl_texture, pub environment: [f32; 3],
    pub nsubmaterials: u32, pub submaterials: *const rl_material, pub environment_texture: u32, pub use_ats: u32,
}
#[repr(C)] pub struct rl_integrator_desc {
    pub kind: u32, pub min_depth: i32, pub max_depth: i32, pub rr_depth: i32, pub strategy: u32,
    pub single_scattering: u32, pub nb_bsdf_samples: u32, pub nb_light_samples: u32,
    pub ao_max_distance: f32, pub ao_normal_correction: u32,
}
#[repr(C)] pub struct rl_render_opts { pub struct_size: u32, pub spp: u32, pub seed: u64, pub sampler_mode: u32, pub batch_spp: u32, pub material_sort: u32, pub sample_offset: u32 }
#[repr(C)] #[derive(Default)]
pub struct rl_stats {
    pub samples: u64, pub segments: u64, pub shadow_rays: u64, pub shadow_visible: u64, pub hits: u64, pub max_depth_seen: u64,
    pub kernel_launches: u64, pub ms_total: f64, pub ms_raygen: f64, pub ms_trace: f64, pub ms_shade: f64, pub ms_shadow: f64,
    pub ms_accum: f64, pub ms_h2d: f64, pub ms_d2h: f64, pub ms_reduce: f64, pub ms_tail: f64, pub shadow_traced: u64,
    pub launches_trace: u64, pub launches_shade: u64,
}
#[repr(C)] #[derive(Default)]
pub struct rl_bvh_info {
    pub ntris: u32, pub nnodes: u32, pub nleaves: u32, pub max_depth: u32, pub root_min: [f32; 3], pub root_max: [f32; 3],
    pub smem_resident: u32, pub flat_groups: u32, pub flat_pairs: u32, pub flat_singles: u32, pub flat_delta: f32,
}
#[repr(C)] #[derive(Default)]
pub struct rl_layout_info {
    pub ray_bytes: u32, pub hit_bytes: u32, pub state_bytes: u32, pub shadow_bytes: u32, pub accum_bytes: u32, pub max_paths_in_flight: u64,
}

#[link(name = "rl_b200")]
extern "C" {
    pub fn rl_create(out: *mut *mut rl_ctx, device: c_int, nranks: c_int, rank: c_int, nccl_unique_id: *const c_void) -> c_int;
    pub fn rl_destroy(ctx: *mut rl_ctx);
    pub fn rl_nccl_unique_id(out_128_bytes: *mut c_void) -> c_int;
    pub fn rl_last_error(ctx: *const rl_ctx) -> *const c_char;
    pub fn rl_abi_version() -> c_int;
    pub fn rl_host_alloc(bytes: usize) -> *mut c_void;
    pub fn rl_host_free(p: *mut c_void);
    pub fn rl_set_profiling(ctx: *mut rl_ctx, on: c_int) -> c_int;
    pub fn rl_scene_create(ctx: *mut rl_ctx, desc: *const rl_scene_desc, out: *mut *mut rl_scene) -> c_int;
    pub fn rl_scene_destroy(ctx: *mut rl_ctx, scene: *mut rl_scene);
    pub fn rl_scene_bvh_info(ctx: *mut rl_ctx, scene: *const rl_scene, out: *mut rl_bvh_info) -> c_int;
    pub fn rl_render(ctx: *mut rl_ctx, scene: *mut rl_scene, integrator: *const rl_integrator_desc,
                     opts: *const rl_render_opts, out_rgb: *mut f32, stats: *mut rl_stats) -> c_int;
    pub fn rl_render_device(ctx: *mut rl_ctx, scene: *mut rl_scene, integrator: *const rl_integrator_desc,
                            opts: *const rl_render_opts, out_rgb_device: *mut f32, stats: *mut rl_stats) -> c_int;
    pub fn rl_trace(ctx: *mut rl_ctx, scene: *mut rl_scene, n: usize, o: *const f32, d: *const f32, prim: *mut u32, tuv: *mut f32) -> c_int;
    pub fn rl_visible(ctx: *mut rl_ctx, scene: *mut rl_scene, n: usize, p0: *const f32, p1: *const f32, out: *mut u8) -> c_int;
    pub fn rl_primary_hits(ctx: *mut rl_ctx, scene: *mut rl_scene, prim: *mut u32, tuv: *mut f32) -> c_int;
    pub fn rl_layout(ctx: *mut rl_ctx, out: *mut rl_layout_info) -> c_int;
}

use crate::integrators::ao::IntegratorAO;
use crate::integrators::direct::IntegratorDirect;
use crate::integrators::explicit::path::{IntegratorPathTracing, IntegratorPathTracingStrategies};
use crate::integrators::{BufferCollection, Integrator};
use crate::{accel::Acceleration, samplers::Sampler, scene::Scene, structure::Color};
use cgmath::{Matrix4, Point2};

fn opt(v: Option<u32>) -> i32 { v.map_or(-1, |x| x as i32) }
unsafe fn last_error(ctx: *const rl_ctx) -> String { CStr::from_ptr(rl_last_error(ctx)).to_string_lossy().into_owned() }
fn col(c: &Color) -> [f32; 3] { [c.r, c.g, c.b] }
fn mat16(m: &Matrix4<f32>) -> [f32; 16] { *AsRef::<[f32; 16]>::as_ref(m) } // cgmath matrices are column-major, like the ABI

/// Host arrays the flat description points into; they live until `rl_scene_create` has returned (the library copies everything).
#[derive(Default)]
pub struct Flat {
    p: Vec<Vec<f32>>, n: Vec<Vec<f32>>, uv: Vec<Vec<f32>>, idx: Vec<Vec<u32>>, texels: Vec<Vec<f32>>,
    meshes: Vec<rl_mesh_desc>, lights: Vec<rl_light_desc>, textures: Vec<rl_texture>, submaterials: Vec<rl_material>,
}
impl Flat {
    /// BSDFColor -> (constant colour, 0) or (black, 1 + texture index); `describe()` of a BSDF calls this for each of its colour slots.
    pub fn color_slot(&mut self, c: &crate::bsdfs::BSDFColor) -> ([f32; 3], u32) {
        use crate::bsdfs::BSDFColor::*;
        let tex = match c {
            Constant(v) => return (col(v), 0),
            Bitmap { img } => {
                self.texels.push(img.colors.iter().flat_map(|c| [c.r, c.g, c.b]).collect());
                rl_texture { kind: 1, width: img.size.x, height: img.size.y, pixels: self.texels.last().unwrap().as_ptr(), color0: [0.0; 3],
                             color1: [0.0; 3], line_width: 0.0, offset: [0.0; 2], scale: [1.0; 2] }
            }
            Checkerbord { color0, color1, offset, scale } => rl_texture { kind: 2, width: 0, height: 0, pixels: std::ptr::null(), color0: col(color0),
                color1: col(color1), line_width: 0.0, offset: [offset.x, offset.y], scale: [scale.x, scale.y] },
            Grid { color0, color1, line_width, offset, scale } => rl_texture { kind: 3, width: 0, height: 0, pixels: std::ptr::null(), color0: col(color0),
                color1: col(color1), line_width: *line_width, offset: [offset.x, offset.y], scale: [scale.x, scale.y] },
        };
        self.textures.push(tex);
        ([0.0; 3], self.textures.len() as u32)
    }
    /// The image of an `EmissionType::Texture` as a bitmap texture: 1 + texture index (rl_mesh_desc.emission_texture).
    pub fn bitmap(&mut self, img: &crate::structure::Bitmap) -> u32 {
        self.texels.push(img.colors.iter().flat_map(|c| [c.r, c.g, c.b]).collect());
        self.textures.push(rl_texture { kind: 1, width: img.size.x, height: img.size.y, pixels: self.texels.last().unwrap().as_ptr(), color0: [0.0; 3],
                                        color1: [0.0; 3], line_width: 0.0, offset: [0.0; 2], scale: [1.0; 2] });
        self.textures.len() as u32
    }
}

/// Scene -> flat description (scene.rs:16-30, geometry.rs:107-119).  Panics where the reference offers something outside
/// the GPU path (media, BSDFs without `describe`), like the reference panics on
/// unsupported input (scene_loader.rs:40-43).
fn flatten(scene: &Scene) -> (Flat, rl_scene_desc) {
    assert!(scene.volume.is_none(), "scene.volume must be None on the GPU path");
    let mut f = Flat::default();
    for m in &scene.meshes {
        f.p.push(m.vertices.iter().flat_map(|v| [v.x, v.y, v.z]).collect());
        f.n.push(m.normals.as_ref().map_or(vec![], |ns| ns.iter().flat_map(|v| [v.x, v.y, v.z]).collect()));
        f.uv.push(m.uv.as_ref().map_or(vec![], |uv| uv.iter().flat_map(|v| [v.x, v.y]).collect()));
        f.idx.push(m.indices.iter().flat_map(|i| [i.x as u32, i.y as u32, i.z as u32]).collect());
    }
    for (k, m) in scene.meshes.iter().enumerate() {
        let mat = m.bsdf.describe(&mut f).unwrap_or_else(|| panic!("mesh {}: this BSDF has no GPU description", m.name));
        let (kind, e, etex) = match &m.emission {
            crate::geometry::EmissionType::Zero => (0, Color::zero(), 0),
            crate::geometry::EmissionType::Color { v } => (1, *v, 0),
            crate::geometry::EmissionType::HSV { scale } => (2, Color::value(*scale), 0), // -x hvs-light (cli.rs:419-420)
            crate::geometry::EmissionType::Texture { scale, img } => (3, Color::value(*scale), f.bitmap(img)), // -x texture-light: 1 + texture index
        };
        let (n, uv) = (&f.n[k], &f.uv[k]);
        f.meshes.push(rl_mesh_desc {
            p: f.p[k].as_ptr(), nverts: m.vertices.len() as u32, idx: f.idx[k].as_ptr(), ntris: m.indices.len() as u32,
            n: if n.is_empty() { std::ptr::null() } else { n.as_ptr() }, uv: if uv.is_empty() { std::ptr::null() } else { uv.as_ptr() },
            mat, emission_kind: kind, emission: col(&e), emission_texture: etex,
        });
    }
    // Non-mesh emitters in the order of Scene.emitters (scene.rs:85-96): mesh lights and the environment are rebuilt by the
    // library from the meshes / `environment`; every other emitter must describe itself (PointEmitter, DirectionalLight).
    if let Some(crate::scene::EmittersState::Build(sampler)) = &scene.emitters {
        for e in &sampler.emitters {
            if e.is_surface() || e.is_environment() { continue; }
            f.lights.push(e.describe().expect("this emitter has no GPU description (PointNormalEmitter?)"));
        }
    }
    let (has_env, env, env_tex) = match scene.emitter_environment.as_ref().map(|e| &e.luminance) {
        None => (0, [0.0; 3], 0),
        Some(crate::emitter::EnvironmentLightColor::Constant(c)) => (1, col(c), 0),
        // the library rebuilds the Distribution2D from the image (EnvironmentLightColor::new_texture, emitter.rs:341-353)
        Some(crate::emitter::EnvironmentLightColor::Texture { image, .. }) => {
            f.texels.push(image.colors.iter().flat_map(|c| [c.r, c.g, c.b]).collect());
            f.textures.push(rl_texture { kind: 1, width: image.size.x, height: image.size.y, pixels: f.texels.last().unwrap().as_ptr(), color0: [0.0; 3],
                                         color1: [0.0; 3], line_width: 0.0, offset: [0.0; 2], scale: [1.0; 2] });
            (2, [0.0; 3], f.textures.len() as u32)
        }
    };
    let cam = &scene.camera;
    let desc = rl_scene_desc {
        nmeshes: f.meshes.len() as u32, meshes: f.meshes.as_ptr(),
        camera: rl_camera_desc { width: cam.size().x, height: cam.size().y, sample_to_camera: mat16(cam.sample_to_camera()), to_world: mat16(cam.to_world()) },
        has_volume: 0, has_environment: has_env, nlights: f.lights.len() as u32, lights: f.lights.as_ptr(),
        ntextures: f.textures.len() as u32, textures: f.textures.as_ptr(), environment: env,
        nsubmaterials: f.submaterials.len() as u32, submaterials: f.submaterials.as_ptr(), environment_texture: env_tex,
        use_ats: matches!(&scene.emitters, Some(crate::scene::EmittersState::Build(s)) if s.ats.is_some()) as u32, // build_emitters(true): `-x ats`
    };
    (f, desc)
}

/// One context + one resident device scene per process, rebuilt when `compute()` sees another `Scene`.
struct Resident { ctx: *mut rl_ctx, dev: *mut rl_scene, scene_addr: usize, seed: u64, passes: u32 }
unsafe impl Send for Resident {}
impl Drop for Resident {
    fn drop(&mut self) { unsafe { rl_scene_destroy(self.ctx, self.dev); rl_destroy(self.ctx); } }
}
static RESIDENT: Mutex<Option<Resident>> = Mutex::new(None);

fn render(scene: &Scene, integ: rl_integrator_desc, seed: u64) -> BufferCollection {
    unsafe {
        assert_eq!(rl_abi_version(), RL_B200_ABI_VERSION, "librl_b200.so and this binding disagree on the ABI");
        let mut guard = RESIDENT.lock().unwrap();
        let addr = scene as *const Scene as usize;
        if guard.as_ref().map_or(true, |r| r.scene_addr != addr) {
            *guard = None; // drops the previous context and scene
            let mut ctx = std::ptr::null_mut();
            if rl_create(&mut ctx, 0, 1, 0, std::ptr::null()) != 0 { panic!("rl_create: {}", last_error(std::ptr::null())); }
            let (_keep, desc) = flatten(scene);
            let mut dev = std::ptr::null_mut();
            if rl_scene_create(ctx, &desc, &mut dev) != 0 { let e = last_error(ctx); rl_destroy(ctx); panic!("rl_scene_create: {}", e); }
            *guard = Some(Resident { ctx, dev, scene_addr: addr, seed, passes: 0 });
        }
        let r = guard.as_mut().unwrap();
        if r.seed != seed { r.seed = seed; r.passes = 0; }
        let size = *scene.camera.size();
        let spp = scene.nb_samples as u32;
        assert_ne!(spp, 0); // integrators/mod.rs:410
        let mut rgb = vec![0f32; (size.x * size.y * 3) as usize];
        // pass p renders samples [p * spp, (p + 1) * spp): an averaging wrapper never sees the same sample twice
        let opts = rl_render_opts { struct_size: std::mem::size_of::<rl_render_opts>() as u32, spp, seed, sampler_mode: 1, batch_spp: 0,
                                    material_sort: 2, sample_offset: r.passes.wrapping_mul(spp) };
        let mut st = rl_stats::default();
        if rl_render(r.ctx, r.dev, &integ, &opts, rgb.as_mut_ptr(), &mut st) != 0 { panic!("rl_render: {}", last_error(r.ctx)); }
        r.passes += 1;
        log::info!("Elapsed Integrator: {} ms", st.ms_total as u64); // same log line as integrators/mod.rs:334
        let mut img = BufferCollection::new(Point2::new(0, 0), size, &["primal".to_string()]);
        for y in 0..size.y { for x in 0..size.x {
            let i = ((y * size.x + x) * 3) as usize;
            img.accumulate(Point2::new(x, y), Color::new(rgb[i], rgb[i + 1], rgb[i + 2]), "primal");
        } }
        img
    }
}

#[cfg(feature = "b200")]
impl Integrator for IntegratorPathTracing {
    fn compute(&mut self, sampler: &mut dyn Sampler, _accel: &dyn Acceleration, scene: &Scene) -> BufferCollection {
        let strategy = match self.strategy { IntegratorPathTracingStrategies::All => 0, IntegratorPathTracingStrategies::BSDF => 1,
                                             IntegratorPathTracingStrategies::Emitter => 2 };
        render(scene, rl_integrator_desc { kind: 0, min_depth: opt(self.min_depth), max_depth: opt(self.max_depth), rr_depth: opt(self.rr_depth),
            strategy, single_scattering: self.single_scattering as u32, nb_bsdf_samples: 1, nb_light_samples: 1,
            ao_max_distance: -1.0, ao_normal_correction: 0 }, sampler.seed())
    }
}
#[cfg(feature = "b200")]
impl Integrator for IntegratorDirect {
    fn compute(&mut self, sampler: &mut dyn Sampler, _accel: &dyn Acceleration, scene: &Scene) -> BufferCollection {
        render(scene, rl_integrator_desc { kind: 1, min_depth: 0, max_depth: -1, rr_depth: 0, strategy: 0, single_scattering: 0,
            nb_bsdf_samples: self.nb_bsdf_samples as u32, nb_light_samples: self.nb_light_samples as u32,
            ao_max_distance: -1.0, ao_normal_correction: 0 }, sampler.seed())
    }
}
#[cfg(feature = "b200")]
impl Integrator for IntegratorAO {
    fn compute(&mut self, sampler: &mut dyn Sampler, _accel: &dyn Acceleration, scene: &Scene) -> BufferCollection {
        render(scene, rl_integrator_desc { kind: 2, min_depth: 0, max_depth: -1, rr_depth: 0, strategy: 0, single_scattering: 0,
            nb_bsdf_samples: 1, nb_light_samples: 0, ao_max_distance: self.max_distance.unwrap_or(-1.0),
            ao_normal_correction: self.normal_correction as u32 }, sampler.seed())
    }
}

/// The additions to the reference's own files, as the code a maintainer pastes in (each block names its file).
mod patch {
    use super::{col, rl_light_desc, rl_material, Flat};
    use crate::bsdfs::distribution::{MicrofacetDistributionBSDF, MicrofacetType};
    use crate::bsdfs::{diffuse::BSDFDiffuse, glass::BSDFGlass, metal::BSDFMetal, phong::BSDFPhong, substrate::BSDFSubstrate, BSDFColor};
    use crate::structure::Color;

    // ---- src/bsdfs/mod.rs, inside `pub trait BSDF`:
    //     /// Flat description for the GPU backend; None = not supported there (BSDFBlend).
    //     fn describe(&self, _flat: &mut crate::b200::Flat) -> Option<crate::b200::rl_material> { None }
    fn blank(kind: u32) -> rl_material {
        rl_material { kind, kd: [0.0; 3], ks: [0.0; 3], exponent: 0.0, weight_specular: 0.0, kt: [0.0; 3], eta: [0.0; 3], k: [0.0; 3],
                      ior: 1.0, alpha: 0.0, microfacet: 0, kd_texture: 0, ks_texture: 0, kt_texture: 0, eta_texture: 0, k_texture: 0, blend_a: 0, blend_b: 0, blend_weight: 0.0 }
    }
    fn microfacet(d: &Option<MicrofacetDistributionBSDF>) -> (u32, f32) {
        match d {
            None => (0, 0.0),
            Some(d) => {
                assert_eq!(d.alpha_u, d.alpha_v); // distribution.rs:62
                (match d.microfacet_type { MicrofacetType::GGX => 1, MicrofacetType::Beckmann => 2 }, d.alpha_u)
            }
        }
    }
    // ---- src/bsdfs/diffuse.rs, inside `impl BSDF for BSDFDiffuse`:
    pub fn describe_diffuse(b: &BSDFDiffuse, flat: &mut Flat) -> Option<rl_material> {
        let (kd, kd_texture) = flat.color_slot(&b.diffuse);
        Some(rl_material { kd, kd_texture, ..blank(0) })
    }
    // ---- src/bsdfs/phong.rs, inside `impl BSDF for BSDFPhong`:
    pub fn describe_phong(b: &BSDFPhong, flat: &mut Flat) -> Option<rl_material> {
        let (kd, kd_texture) = flat.color_slot(&b.diffuse);
        let (ks, ks_texture) = flat.color_slot(&b.specular);
        Some(rl_material { kd, kd_texture, ks, ks_texture, exponent: b.exponent, weight_specular: b.weight_specular, ..blank(1) })
    }
    // ---- src/bsdfs/metal.rs, inside `impl BSDF for BSDFMetal`:
    pub fn describe_metal(b: &BSDFMetal, flat: &mut Flat) -> Option<rl_material> {
        let (microfacet, alpha) = microfacet(&b.distribution);
        let (ks, ks_texture) = flat.color_slot(&b.specular);
        let (eta, eta_texture) = flat.color_slot(&b.eta);
        let (k, k_texture) = flat.color_slot(&b.k);
        Some(rl_material { ks, ks_texture, eta, eta_texture, k, k_texture, microfacet, alpha, ..blank(2) })
    }
    // ---- src/bsdfs/glass.rs, inside `impl BSDF for BSDFGlass`:
    pub fn describe_glass(b: &BSDFGlass, flat: &mut Flat) -> Option<rl_material> {
        let (ks, ks_texture) = flat.color_slot(&b.specular_reflectance);
        let (kt, kt_texture) = flat.color_slot(&b.specular_transmittance);
        Some(rl_material { ks, ks_texture, kt, kt_texture, ior: b.eta, ..blank(3) })
    }
    // ---- src/bsdfs/substrate.rs, inside `impl BSDF for BSDFSubstrate`:
    pub fn describe_substrate(b: &BSDFSubstrate, flat: &mut Flat) -> Option<rl_material> {
        let (kd, kd_texture) = flat.color_slot(&b.diffuse);
        let (ks, ks_texture) = flat.color_slot(&b.specular);
        let (microfacet, alpha) = microfacet(&b.distribution);
        Some(rl_material { kd, kd_texture, ks, ks_texture, microfacet, alpha, ..blank(4) })
    }
    // ---- src/bsdfs/blend.rs, inside `impl BSDF for BSDFBlend` (both parts must describe themselves and be rough, blend.rs:17):
    pub fn describe_blend(b: &crate::bsdfs::blend::BSDFBlend, flat: &mut Flat) -> Option<rl_material> {
        let (a, c) = (b.bsdf1.describe(flat)?, b.bsdf2.describe(flat)?);
        flat.submaterials.push(a);
        flat.submaterials.push(c);
        let n = flat.submaterials.len() as u32;
        Some(rl_material { blend_a: n - 1, blend_b: n, blend_weight: b.weight, ..blank(5) })
    }
    // (each `impl BSDF for X` gains `fn describe(&self, flat: &mut Flat) -> Option<rl_material> { crate::b200::patch::describe_x(self, flat) }`)

    // ---- src/emitter.rs, inside `pub trait Emitter`:
    //     fn describe(&self) -> Option<crate::b200::rl_light_desc> { None }
    //     fn is_environment(&self) -> bool { false }        // EnvironmentLight returns true
    // ---- inside `impl Emitter for PointEmitter`:
    pub fn describe_point(e: &crate::emitter::PointEmitter) -> Option<rl_light_desc> {
        Some(rl_light_desc { kind: 0, intensity: col(&e.intensity), v: [e.position.x, e.position.y, e.position.z] })
    }
    // ---- inside `impl Emitter for DirectionalLight` (its bounding sphere is rebuilt by the library, scene.rs:54-60):
    pub fn describe_directional(e: &crate::emitter::DirectionalLight) -> Option<rl_light_desc> {
        Some(rl_light_desc { kind: 1, intensity: col(&e.intensity), v: [e.direction.x, e.direction.y, e.direction.z] })
    }

    // ---- src/camera.rs, inside `impl Camera` (the matrices built by Camera::new, camera.rs:31-67, are private fields):
    //     pub fn sample_to_camera(&self) -> &Matrix4<f32> { &self.sample_to_camera }
    //     pub fn to_world(&self) -> &Matrix4<f32> { &self.to_world }

    // ---- src/samplers/mod.rs, inside `pub trait Sampler`:
    //     /// Seed of `-r independent:<seed>` (cli.rs:886-890); the GPU backend keys its counter streams on it.
    //     fn seed(&self) -> u64 { 0 }
    // ---- src/samplers/independent.rs: `pub struct IndependentSampler { pub rnd: SmallRng, pub seed: u64 }`, set where the CLI
    //      builds it (`IndependentSampler { rnd: SmallRng::seed_from_u64(s), seed: s }`), and `fn seed(&self) -> u64 { self.seed }`.
    #[allow(unused)]
    fn _unused(_: Color) {}
}
