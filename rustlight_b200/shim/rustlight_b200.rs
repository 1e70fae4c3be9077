//! rustlight_b200.rs -- the binding a rustlight maintainer would add (UNCOMPILED here: this image
//! has no Rust toolchain).  It declares the C ABI of include/rl_b200.h and
//! implements `Integrator` for the two integrators by flattening `Scene` into `rl_scene_desc`,
//! calling `rl_render` and wrapping the result into a `BufferCollection` -- replacing the
//! one-line `compute_mc(self, sampler, accel, scene)` bodies of
//! src/integrators/explicit/path.rs:187-196 and src/integrators/direct.rs:10-19.
//! The `accel` argument is ignored: the GPU library builds its own LBVH.
#![allow(non_camel_case_types, dead_code)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct rl_ctx { _p: [u8; 0] }
#[repr(C)] pub struct rl_scene { _p: [u8; 0] }

#[repr(C)] #[derive(Clone, Copy)]
pub struct rl_material {
    pub kind: u32, pub kd: [f32; 3], pub ks: [f32; 3], pub exponent: f32, pub weight_specular: f32,
    pub kt: [f32; 3], pub eta: [f32; 3], pub k: [f32; 3], pub ior: f32, pub alpha: f32, pub microfacet: u32, pub kd_texture: u32,
}
#[repr(C)] pub struct rl_texture {
    pub kind: u32, pub width: u32, pub height: u32, pub pixels: *const f32, pub color0: [f32; 3], pub color1: [f32; 3],
    pub line_width: f32, pub offset: [f32; 2], pub scale: [f32; 2],
}
#[repr(C)]
pub struct rl_mesh_desc {
    pub p: *const f32, pub nverts: u32, pub idx: *const u32, pub ntris: u32,
    pub n: *const f32, pub uv: *const f32, pub mat: rl_material, pub emission_kind: u32, pub emission: [f32; 3],
}
#[repr(C)] pub struct rl_camera_desc { pub width: u32, pub height: u32, pub sample_to_camera: [f32; 16], pub to_world: [f32; 16] }
#[repr(C)] pub struct rl_light_desc { pub kind: u32, pub intensity: [f32; 3], pub v: [f32; 3] }
#[repr(C)] pub struct rl_scene_desc {
    pub nmeshes: u32, pub meshes: *const rl_mesh_desc, pub camera: rl_camera_desc, pub has_volume: u32, pub has_environment: u32,
    pub nlights: u32, pub lights: *const rl_light_desc, pub ntextures: u32, pub textures: *const rl_texture, pub environment: [f32; 3],
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
    pub ms_accum: f64, pub ms_h2d: f64, pub ms_d2h: f64, pub ms_reduce: f64,
}

#[link(name = "rl_b200")]
extern "C" {
    pub fn rl_create(out: *mut *mut rl_ctx, device: c_int, nranks: c_int, rank: c_int, nccl_unique_id: *const c_void) -> c_int;
    pub fn rl_destroy(ctx: *mut rl_ctx);
    pub fn rl_last_error(ctx: *const rl_ctx) -> *const c_char;
    pub fn rl_scene_create(ctx: *mut rl_ctx, desc: *const rl_scene_desc, out: *mut *mut rl_scene) -> c_int;
    pub fn rl_scene_destroy(ctx: *mut rl_ctx, scene: *mut rl_scene);
    pub fn rl_render(ctx: *mut rl_ctx, scene: *mut rl_scene, integrator: *const rl_integrator_desc,
                     opts: *const rl_render_opts, out_rgb: *mut f32, stats: *mut rl_stats) -> c_int;
    pub fn rl_trace(ctx: *mut rl_ctx, scene: *mut rl_scene, n: usize, o: *const f32, d: *const f32, prim: *mut u32, tuv: *mut f32) -> c_int;
    pub fn rl_visible(ctx: *mut rl_ctx, scene: *mut rl_scene, n: usize, p0: *const f32, p1: *const f32, out: *mut u8) -> c_int;
    pub fn rl_host_alloc(bytes: usize) -> *mut c_void; // optional: pinned buffer for out_rgb
    pub fn rl_host_free(p: *mut c_void);
}

use crate::bsdfs::BSDFType;
use crate::integrators::{BufferCollection, Integrator};
use crate::integrators::direct::IntegratorDirect;
use crate::integrators::explicit::path::{IntegratorPathTracing, IntegratorPathTracingStrategies};
use crate::{accel::Acceleration, samplers::Sampler, scene::Scene, structure::Color};
use cgmath::{Matrix, Point2};

fn opt(v: Option<u32>) -> i32 { v.map_or(-1, |x| x as i32) }

/// Scene -> flat description.  Vectors are kept alive in `Flat` for the duration of the call.
struct Flat { p: Vec<Vec<f32>>, n: Vec<Vec<f32>>, idx: Vec<Vec<u32>>, meshes: Vec<rl_mesh_desc>, lights: Vec<rl_light_desc>, textures: Vec<rl_texture> }
fn flatten(scene: &Scene) -> (Flat, rl_scene_desc) {
    assert!(scene.volume.is_none(), "scene.volume must be None on the GPU path");
    let mut f = Flat { p: vec![], n: vec![], idx: vec![], meshes: vec![], lights: vec![], textures: vec![] };
    for m in &scene.meshes {
        f.p.push(m.vertices.iter().flat_map(|v| [v.x, v.y, v.z]).collect());
        f.n.push(m.normals.as_ref().map_or(vec![], |ns| ns.iter().flat_map(|v| [v.x, v.y, v.z]).collect()));
        f.idx.push(m.indices.iter().flat_map(|i| [i.x as u32, i.y as u32, i.z as u32]).collect());
    }
    for (k, m) in scene.meshes.iter().enumerate() {
        // BSDFDiffuse / BSDFPhong -> rl_material: needs a small `fn describe(&self) -> rl_material`
        // on the BSDF trait (or a downcast); everything else on this path is rejected up front.
        let mat = m.bsdf.describe();
        let (kind, e) = match &m.emission {
            crate::geometry::EmissionType::Zero => (0, Color::zero()),
            crate::geometry::EmissionType::Color { v } => (1, *v),
            _ => panic!("textured emission is outside the GPU path"),
        };
        f.meshes.push(rl_mesh_desc {
            p: f.p[k].as_ptr(), nverts: m.vertices.len() as u32, idx: f.idx[k].as_ptr(), ntris: m.indices.len() as u32,
            n: if f.n[k].is_empty() { std::ptr::null() } else { f.n[k].as_ptr() }, uv: std::ptr::null(),
            mat, emission_kind: kind, emission: [e.r, e.g, e.b],
        });
    }
    let cam = &scene.camera;
    let mut s2c = [0f32; 16]; let mut c2w = [0f32; 16];
    s2c.copy_from_slice(AsRef::<[f32; 16]>::as_ref(cam.sample_to_camera()));   // needs pub accessors in camera.rs
    c2w.copy_from_slice(AsRef::<[f32; 16]>::as_ref(cam.to_world()));
    let desc = rl_scene_desc { nmeshes: f.meshes.len() as u32, meshes: f.meshes.as_ptr(),
        camera: rl_camera_desc { width: cam.size().x, height: cam.size().y, sample_to_camera: s2c, to_world: c2w },
        has_volume: 0, has_environment: scene.emitter_environment.is_some() as u32,
        // PointEmitter / DirectionalLight of EmittersState::Unbuild need a `describe() -> Option<rl_light_desc>` on the Emitter
        // trait; `f.lights` keeps them alive like the mesh vectors
        nlights: f.lights.len() as u32, lights: f.lights.as_ptr(),
        // BSDFColor::{Bitmap, Checkerbord, Grid} on a diffuse slot -> rl_texture + rl_material.kd_texture (describe() fills both)
        ntextures: f.textures.len() as u32, textures: f.textures.as_ptr(),
        // EnvironmentLightColor::Constant(c) -> has_environment = 1 + environment = c; a Texture environment must be rejected
        environment: env_constant(scene) };
    (f, desc)
}

fn render(scene: &Scene, integ: rl_integrator_desc, seed: u64) -> BufferCollection {
    unsafe {
        let mut ctx = std::ptr::null_mut();
        assert_eq!(rl_create(&mut ctx, 0, 1, 0, std::ptr::null()), 0, "{:?}", CStr::from_ptr(rl_last_error(std::ptr::null())));
        let (_keep, desc) = flatten(scene);
        let mut dev = std::ptr::null_mut();
        if rl_scene_create(ctx, &desc, &mut dev) != 0 { panic!("{:?}", CStr::from_ptr(rl_last_error(ctx))); }
        let size = *scene.camera.size();
        let mut rgb = vec![0f32; (size.x * size.y * 3) as usize];
        let opts = rl_render_opts { struct_size: std::mem::size_of::<rl_render_opts>() as u32, spp: scene.nb_samples as u32,
            seed, sampler_mode: 1, batch_spp: 0, material_sort: 0, sample_offset: 0 };
        let mut st = rl_stats::default();
        if rl_render(ctx, dev, &integ, &opts, rgb.as_mut_ptr(), &mut st) != 0 { panic!("{:?}", CStr::from_ptr(rl_last_error(ctx))); }
        info!("Elapsed Integrator: {} ms", st.ms_total as u64);   // same log line as integrators/mod.rs:334
        let mut img = BufferCollection::new(Point2::new(0, 0), size, &["primal".to_string()]);
        for y in 0..size.y { for x in 0..size.x {
            let i = ((y * size.x + x) * 3) as usize;
            img.accumulate(Point2::new(x, y), Color::new(rgb[i], rgb[i + 1], rgb[i + 2]), &"primal".to_string());
        } }
        rl_scene_destroy(ctx, dev);
        rl_destroy(ctx);
        img
    }
}

/// `-r independent:<seed>`: the sampler only contributes its seed (Sampler gains `fn seed(&self) -> u64`).
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
impl Integrator for crate::integrators::ao::IntegratorAO {
    fn compute(&mut self, sampler: &mut dyn Sampler, _accel: &dyn Acceleration, scene: &Scene) -> BufferCollection {
        render(scene, rl_integrator_desc { kind: 2, min_depth: 0, max_depth: -1, rr_depth: 0, strategy: 0, single_scattering: 0,
            nb_bsdf_samples: 1, nb_light_samples: 0, ao_max_distance: self.max_distance.unwrap_or(-1.0),
            ao_normal_correction: self.normal_correction as u32 }, sampler.seed())
    }
}
