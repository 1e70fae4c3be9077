"""Known-answer / analytic tests that pin the oracle (the reference ships none: SURVEY.md F4).

Each test names the reference item whose arithmetic it pins.
"""
import math

import numpy as np
import pytest

from oracle import binding as ob
from rustlight_b200 import _abi
from rustlight_b200.host import material_diffuse, material_phong

F32_MAX = np.finfo(np.float32).max


# ---- Mesh::intersection_tri, geometry.rs:358-410 -------------------------------------------------
V0, V1, V2 = (0, 0, 0), (1, 0, 0), (0, 1, 0)


def test_tri_centre_hit():
    hit, t, u, v, p, n = ob.intersect_tri(V0, V1, V2, (0.25, 0.25, 1), (0, 0, -1))
    assert hit and t == 1.0 and u == 0.25 and v == 0.25
    assert np.array_equal(p, [0.25, 0.25, 0]) and np.array_equal(n, [0, 0, 1])


def test_tri_two_sided():
    # the plane test has no back-face culling: the same triangle is hit from below
    hit, t, u, v, _, n = ob.intersect_tri(V0, V1, V2, (0.25, 0.25, -2), (0, 0, 1))
    assert hit and t == 2.0 and (u, v) == (0.25, 0.25) and np.array_equal(n, [0, 0, 1])


@pytest.mark.parametrize("o", [(0.5, 0.5, 1), (0.0, 0.0, 1), (1.0, 0.0, 1), (0.0, 1.0, 1), (0.5, 0.0, 1), (0.0, 0.5, 1)])
def test_tri_edges_and_vertices_are_inside(o):
    # u,v in [0,1] and u+v<=1 are inclusive (geometry.rs:391-394)
    hit, t, u, v, _, _ = ob.intersect_tri(V0, V1, V2, o, (0, 0, -1))
    assert hit and t == 1.0 and abs(u - o[0]) < 1e-7 and abs(v - o[1]) < 1e-7


@pytest.mark.parametrize("o", [(0.51, 0.5, 1), (-1e-3, 0.5, 1), (0.5, -1e-3, 1), (2, 2, 1)])
def test_tri_outside(o):
    assert not ob.intersect_tri(V0, V1, V2, o, (0, 0, -1))[0]


def test_tri_parallel_and_behind():
    assert not ob.intersect_tri(V0, V1, V2, (0.2, 0.2, 1), (1, 0, 0))[0]       # denom == 0
    assert not ob.intersect_tri(V0, V1, V2, (0.2, 0.2, 1), (0, 0, 1))[0]       # t < 0


def test_tri_self_intersection_epsilon():
    # accepted only when t > 1e-5 (geometry.rs:398) ...
    assert not ob.intersect_tri(V0, V1, V2, (0.2, 0.2, 5e-6), (0, 0, -1))[0]
    assert ob.intersect_tri(V0, V1, V2, (0.2, 0.2, 2e-5), (0, 0, -1))[0]
    # ... and only when closer than the current its.t (strict)
    assert not ob.intersect_tri(V0, V1, V2, (0.2, 0.2, 1), (0, 0, -1), t_max=1.0)[0]
    assert ob.intersect_tri(V0, V1, V2, (0.2, 0.2, 1), (0, 0, -1), t_max=1.0000001)[0]


def test_tri_degenerate_is_never_hit():
    assert not ob.intersect_tri((0, 0, 0), (1, 0, 0), (2, 0, 0), (0.5, 0, 1), (0, 0, -1))[0]


# ---- AABB::intersect, structure.rs:849-869 ----------------------------------------------------------
def test_aabb_slab():
    lo, hi = (-1, -1, -1), (1, 1, 1)
    assert ob.aabb_intersect(lo, hi, (0, 0, 5), (0, 0, -1)) == (True, 4.0)
    assert ob.aabb_intersect(lo, hi, (0, 0, 0), (0, 0, -1)) == (True, pytest.approx(1e-4))  # inside: returns tnear
    assert not ob.aabb_intersect(lo, hi, (0, 0, 5), (0, 0, 1))[0]
    assert not ob.aabb_intersect(lo, hi, (2, 0, 5), (0, 0, -1))[0]       # parallel outside: inf slabs
    assert not ob.aabb_intersect(lo, hi, (0, 0, 5), (0, 0, -1), tfar=3.0)[0]
    # grazing the face exactly: t_max <= t_min rejects (structure.rs:863)
    assert not ob.aabb_intersect(lo, hi, (0, 1, 5), (0, 1, 0))[0]


# ---- Frame, math.rs:357-384 ----------------------------------------------------------------------------
@pytest.mark.parametrize("n", [(0, 0, 1), (0, 0, -1), (1, 0, 0), (0, 1, 0), (0.6, 0, 0.8), (-0.36, 0.48, -0.8)])
def test_frame_orthonormal(n):
    f = ob.frame(n).astype(np.float64)
    assert np.allclose(f @ f.T, np.eye(3), atol=2e-7)
    assert np.array_equal(f[2], np.float32(n))
    assert np.linalg.det(f) == pytest.approx(1.0, abs=1e-6)


def test_frame_signum_of_negative_zero():
    # f32::signum(-0.0) == -1: the two poles use different branches
    a, b = ob.frame((0, 0, 1.0)), ob.frame((0, 0, -1.0))
    assert np.array_equal(a[0], [1, 0, 0]) and np.array_equal(b[0], [1, 0, 0]) and b[1][1] == -1.0


# ---- sampling, math.rs:37-65, 388-394 ---------------------------------------------------------------------
@pytest.mark.parametrize("mode", [ob.MATH_LIBM, ob.MATH_SPEC])
def test_cosine_sample_hemisphere(mode):
    assert np.array_equal(ob.cosine_sample_hemisphere(0.5, 0.5, mode), [0, 0, 1])  # degenerate centre
    rng = np.random.default_rng(1)
    zs = []
    for u0, u1 in rng.random((2000, 2)):
        d = ob.cosine_sample_hemisphere(float(u0), float(u1), mode)
        assert d[2] >= 0 and abs(np.linalg.norm(d.astype(np.float64)) - 1) < 1e-6
        zs.append(d[2])
    assert np.mean(zs) == pytest.approx(2 / 3, abs=0.02)  # E[cos] under the cos/pi density


def test_cosine_sample_hemisphere_corners():
    for u in [(0, 0), (1 - 2**-24, 0), (0, 1 - 2**-24), (1 - 2**-24, 1 - 2**-24), (0.75, 0.25)]:
        a = ob.cosine_sample_hemisphere(*u, ob.MATH_LIBM).astype(np.float64)
        b = ob.cosine_sample_hemisphere(*u, ob.MATH_SPEC).astype(np.float64)
        assert np.allclose(a[:2], b[:2], atol=3e-7)
        assert abs(a[2] - b[2]) <= 5e-4  # z = sqrt(1 - x^2 - y^2) near the rim magnifies a 1-ulp difference of x, y


def test_uniform_sample_triangle():
    assert np.array_equal(ob.uniform_sample_triangle(0.0, 0.3), [1, 0])
    b = ob.uniform_sample_triangle(0.25, 0.5)
    assert np.array_equal(b, [0.5, 0.25])
    rng = np.random.default_rng(2)
    for u0, u1 in rng.random((500, 2)):
        b = ob.uniform_sample_triangle(float(u0), float(u1))
        assert b[0] >= 0 and b[1] >= 0 and b[0] + b[1] <= 1 + 1e-7


# ---- Distribution1D, math.rs:398-487 -------------------------------------------------------------------------
def test_dist1d_normalize_quirks():
    cdf, func_int = ob.dist1d_normalize([1.0, 3.0])
    # increments are e/n (math.rs:426): func_int = (1+3)/2, total() = func_int * n = 4
    assert func_int == 2.0 and np.array_equal(cdf, [0, 0.25, 1.0])
    cdf, func_int = ob.dist1d_normalize([0.0, 0.0])
    assert func_int == 0.0 and np.array_equal(cdf, [0, 0, 1.0])  # last entry forced to 1 (math.rs:434)


def test_dist1d_sample_discrete():
    cdf = [0, 0.25, 1.0]
    assert ob.dist1d_sample_discrete(cdf, 0.0) == 0          # Ok(0)
    assert ob.dist1d_sample_discrete(cdf, 0.2499) == 0       # Err(1) -> 0
    assert ob.dist1d_sample_discrete(cdf, 0.25) == 1         # Ok(1)
    assert ob.dist1d_sample_discrete(cdf, 0.9999) == 1       # Err(2) -> 1


# ---- mis_weight (power heuristic), integrators/mod.rs:462-478 ------------------------------------------------------
def test_mis_weight():
    assert ob.mis_weight(1.0, 1.0) == 0.5
    assert ob.mis_weight(3.0, 4.0) == np.float32(9.0) / np.float32(25.0)
    assert ob.mis_weight(0.0, 1.0) == 0.0
    assert ob.mis_weight(float("inf"), 1.0) == 0.0 and ob.mis_weight(1.0, float("nan")) == 0.0
    assert ob.mis_weight(1e30, 1.0) == 0.0  # a*a overflows -> inf/inf -> not finite -> 0


# ---- samplers: rand 0.8.5 SmallRng == xoshiro256++ (published algorithm) ------------------------------------------------
M64 = (1 << 64) - 1


def _rotl(x, k):
    return ((x << k) | (x >> (64 - k))) & M64


def _xoshiro_py(s):
    s = list(s)
    r = (_rotl((s[0] + s[3]) & M64, 23) + s[0]) & M64
    t = (s[1] << 17) & M64
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = _rotl(s[3], 45)
    return r, s


def test_xoshiro256pp_reference_vector():
    # first outputs for state (1,2,3,4): the vector used by the xoshiro256++ reference tests
    expect = [41943041, 58720359, 3588806011781223, 3591011842654386, 9228616714210784205,
              9973669472204895162, 14011001112246962877, 12406186145184390807, 15849039046786891736, 10450023813501588000]
    st = [1, 2, 3, 4]
    stp = [1, 2, 3, 4]
    for e in expect:
        r, st = ob.xoshiro_next_u64(st)
        rp, stp = _xoshiro_py(stp)
        assert r == rp == e and st == stp


def _mix64(z):
    z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & M64
    z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & M64
    return z ^ (z >> 31)


def test_counter_sampler_matches_its_definition():
    # key = mix(mix(seed+PHI) ^ (pixel<<32|sample)); draw n = a 24-bit half of z = mix(key + ((n+1)//2)*PHI):
    # bits 40..63 for odd n, bits 16..39 for even n; times 2^-24
    PHI = 0x9e3779b97f4a7c15
    for seed, pixel, sample in [(0, 0, 0), (7, 123456, 99), (2**63 + 5, 2**32 - 1, 2**32 - 1)]:
        key = _mix64(_mix64((seed + PHI) & M64) ^ ((pixel << 32) | sample))
        def draw(n):
            z = _mix64((key + ((n + 1) // 2) * PHI) & M64)
            return (z >> 40) if n & 1 else ((z >> 16) & 0xFFFFFF)
        exp = [np.float32(draw(n)) * np.float32(2.0**-24) for n in range(1, 9)]
        got = ob.sampler_counter(seed, pixel, sample, 8)
        assert np.array_equal(got, np.array(exp, np.float32))
        assert (got >= 0).all() and (got < 1).all()


def test_splitmix64_known_answer():
    # SplitMix64 reference outputs for seed 1234567 (Vigna's splitmix64.c)
    PHI = 0x9e3779b97f4a7c15
    s = 1234567
    out = []
    for _ in range(3):
        s = (s + PHI) & M64
        out.append(_mix64(s))
    assert out == [6457827717110365317, 3203168211198807973, 9817491932198370423]


@pytest.mark.parametrize("seeding", [ob.SEED_PCG32, ob.SEED_SPLITMIX64])
def test_block_stream_is_uniform_and_blocks_differ(seeding):
    a = ob.sampler_block_stream(0, 0, 4096, seeding)
    b = ob.sampler_block_stream(0, 1, 4096, seeding)
    assert (a >= 0).all() and (a < 1).all() and not np.array_equal(a, b)
    assert a.mean() == pytest.approx(0.5, abs=0.02) and a.var() == pytest.approx(1 / 12, abs=0.01)
    # 24-bit lattice (rand's Standard f32)
    assert np.array_equal(a * np.float32(2**24), np.round(a * np.float32(2**24)))


# ---- spec transcendental functions vs libm ----------------------------------------------------------------------------
def test_spec_sincos_is_within_one_and_a_half_ulp():
    """f32 Cody-Waite + Cephes kernels: |error| <= 9e-8 absolute and <= 1.5 ulp of the exact value on
    the ranges the callers use (concentric disk: [-pi/4, 3pi/4]; Phong / microfacet / sphere: [0, 2 pi])."""
    xs = np.concatenate([np.linspace(-0.79, 2.36, 4001), np.linspace(0, 6.2832, 4001)]).astype(np.float32)
    bad = 0
    for x in xs:
        s, c = ob.spec_sincos(float(x))
        for got, exact in ((s, math.sin(float(x))), (c, math.cos(float(x)))):
            e32 = np.float32(exact)
            bad += got != e32
            assert abs(got - exact) <= 9e-8
            assert abs(got - exact) <= 1.5 * float(np.spacing(np.abs(e32))) or abs(exact) < 1e-6
    assert bad < 0.25 * 2 * len(xs)  # most values are the correctly rounded ones


def test_spec_powf():
    rng = np.random.default_rng(3)
    for x, y in zip(rng.random(3000).astype(np.float32), rng.uniform(0.01, 120, 3000).astype(np.float32)):
        got = ob.spec_powf(float(x), float(y))
        exp = np.float32(math.pow(float(x), float(y)))
        assert got == exp or abs(got - float(exp)) <= 1.2e-7 * abs(float(exp))
    assert ob.spec_powf(0.0, 2.0) == 0.0 and ob.spec_powf(1.0, 50.0) == 1.0 and ob.spec_powf(0.5, 0.0) == 1.0


# ---- BSDF estimator identities: diffuse.rs, phong.rs -----------------------------------------------------------------------
MATS = [material_diffuse((0.7, 0.5, 0.3)), material_phong((0.3, 0.35, 0.2), (0.3, 0.3, 0.3), 50.0),
        material_phong((0.0, 0.0, 0.0), (0.9, 0.9, 0.9), 5.0)]


@pytest.mark.parametrize("mat", MATS)
@pytest.mark.parametrize("mode", [ob.MATH_LIBM, ob.MATH_SPEC])
def test_bsdf_weight_is_eval_over_pdf(mat, mode):
    rng = np.random.default_rng(4)
    wi = np.float32([0.3, -0.2, math.sqrt(1 - 0.13)])
    n_ok = 0
    for s0, s1 in rng.random((400, 2)):
        ok, w, d, pdf = ob.bsdf_sample(mat, wi, float(s0), float(s1), mode)
        if not ok:
            continue
        n_ok += 1
        assert d[2] > 0 and pdf > 0
        assert pdf == pytest.approx(ob.bsdf_pdf(mat, wi, d, mode), rel=1e-6)
        ev = ob.bsdf_eval(mat, wi, d, mode)
        if mat.kind == _abi.RL_BSDF_DIFFUSE:
            assert np.array_equal(w, np.float32(list(mat.kd)))  # weight == albedo (diffuse.rs:23)
        else:
            assert np.allclose(w, ev / np.float32(pdf), rtol=1e-6)
    assert n_ok > 300


@pytest.mark.parametrize("mat", MATS)
def test_bsdf_pdf_integrates_to_one(mat):
    wi = np.float32([0.2, 0.1, math.sqrt(1 - 0.05)])
    nt, nph = 400, 400
    th = (np.arange(nt) + 0.5) * (math.pi / 2) / nt
    ph = (np.arange(nph) + 0.5) * (2 * math.pi) / nph
    tot = 0.0
    for t in th:
        st, ct = math.sin(t), math.cos(t)
        row = sum(ob.bsdf_pdf(mat, wi, (st * math.cos(p), st * math.sin(p), ct)) for p in ph[::8]) * 8
        tot += row * st
    tot *= (math.pi / 2 / nt) * (2 * math.pi / nph)
    # Phong's specular lobe leaks below the horizon, so its hemispherical integral is <= 1
    if mat.kind == _abi.RL_BSDF_DIFFUSE:
        assert tot == pytest.approx(1.0, abs=5e-3)
    else:
        assert 0.9 < tot <= 1.0 + 5e-3


def test_bsdf_below_horizon():
    for mat in MATS:
        assert not ob.bsdf_sample(mat, (0, 0, -1), 0.3, 0.3)[0]
        assert ob.bsdf_pdf(mat, (0, 0, 1), (0, 0, -1)) == 0.0
        assert not ob.bsdf_eval(mat, (0, 0, -1), (0, 0, 1)).any()


def test_phong_weight_specular():
    m = material_phong((0.2, 0.2, 0.2), (0.6, 0.6, 0.6), 10.0)
    assert m.weight_specular == pytest.approx(0.75, rel=1e-6)  # lum(ks)/(lum(kd)+lum(ks)), bsdfs/mod.rs:518-523


# ---- outside pins of the third-party arithmetic the oracle restates -----------------------------
def _pcg32_fill(state, nbytes):
    """rand_core 0.6.4 SeedableRng::seed_from_u64 (the default SmallRng 0.8.5 inherits): PCG32 steps, output XSH-RR, little endian."""
    MUL, INC = 6364136223846793005, 11634580027462260723
    out = b""
    while len(out) < nbytes:
        state = (state * MUL + INC) & M64
        xorshifted = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
        rot = state >> 59
        x = ((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & 0xFFFFFFFF
        out += x.to_bytes(4, "little")
    return out[:nbytes]


def _splitmix_fill(state, nbytes):
    """Xoshiro256PlusPlus::seed_from_u64 (rand_xoshiro / rand 0.8.5 xoshiro256plusplus.rs): SplitMix64 outputs, little endian."""
    PHI = 0x9e3779b97f4a7c15
    out = b""
    while len(out) < nbytes:
        state = (state + PHI) & M64
        out += _mix64(state).to_bytes(8, "little")
    return out[:nbytes]


def _seed_from_u64_py(seed, seeding):
    raw = _pcg32_fill(seed, 32) if seeding == ob.SEED_PCG32 else _splitmix_fill(seed, 32)
    return [int.from_bytes(raw[8 * i:8 * i + 8], "little") for i in range(4)]  # Xoshiro256PlusPlus::from_seed: read_u64_into


def test_pcg32_known_answer():
    """PCG32 (XSH-RR 64/32) with rand_core's constants: state 0 -> first outputs, computed by hand from the definition."""
    raw = _pcg32_fill(0, 8)
    # state1 = INC; xorshifted = ((INC >> 18) ^ INC) >> 27; rot = INC >> 59
    INC = 11634580027462260723
    xs = (((INC >> 18) ^ INC) >> 27) & 0xFFFFFFFF
    rot = INC >> 59
    assert int.from_bytes(raw[:4], "little") == ((xs >> rot) | (xs << (32 - rot))) & 0xFFFFFFFF
    assert rot == 20 and len(set(raw)) > 4


@pytest.mark.parametrize("seeding", [ob.SEED_PCG32, ob.SEED_SPLITMIX64])
@pytest.mark.parametrize("seed", [0, 1, 0xDEADBEEFCAFEF00D])
def test_block_stream_against_an_independent_python_implementation(seeding, seed):
    """Sampler mode A end to end (samplers/independent.rs:10-22, integrators/mod.rs:351-374): master = SmallRng::seed_from_u64(seed);
    block k's sampler = SmallRng::seed_from_u64(master.next_u64()) cloned in block order; next() = (next_u64() >> 32 >> 8) * 2^-24.
    Every step re-typed here in Python from the published algorithms (PCG32 fill / SplitMix64 fill, xoshiro256++)."""
    master = _seed_from_u64_py(seed, seeding)
    for block in range(3):
        r, master = _xoshiro_py(master)
        st = _seed_from_u64_py(r, seeding)
        exp = []
        for _ in range(16):
            v, st = _xoshiro_py(st)
            exp.append(np.float32((v >> 32) >> 8) * np.float32(2.0**-24))
        got = ob.sampler_block_stream(seed, block, 16, seeding)
        assert np.array_equal(got, np.array(exp, np.float32)), (seeding, seed, block)


def test_cgmath_perspective_and_invert_against_numpy_f64():
    """Camera::new (camera.rs:31-67) = (S(-1/2, -a/2, 1) T(-1, -1/a, 0) perspective(fov, 1, 1e-2, 1000) S(x_v, 1, -1))^-1 with cgmath 0.18's
    perspective (c0r0 = cot(f/2)/aspect, c1r1 = cot(f/2), c2r2 = (far+near)/(near-far), c2r3 = -1, c3r2 = 2 far near/(near-far)) and
    Matrix4::invert: rebuilt in float64 with numpy.linalg.inv and compared entry by entry."""
    def scale(x, y, z):
        return np.diag([x, y, z, 1.0])

    def trans(x, y, z):
        m = np.eye(4)
        m[:3, 3] = [x, y, z]
        return m

    def perspective(fovy_deg, aspect, near, far):
        f = 1.0 / math.tan(math.radians(fovy_deg) / 2.0)
        m = np.zeros((4, 4))
        m[0, 0], m[1, 1] = f / aspect, f
        m[2, 2], m[2, 3] = (far + near) / (near - far), 2.0 * far * near / (near - far)
        m[3, 2] = -1.0
        return m
    for (w, h, fov, axis, flip) in [(512, 512, 19.5, "y", False), (1920, 1080, 30.0, "y", False), (640, 480, 45.0, "x", True), (64, 128, 60.0, "y", True)]:
        a = w / h
        fovy = fov * a if axis == "y" else fov  # the Fov::Y quirk (camera.rs:41-44): v * aspect, Fov::X(v): v
        c2s = scale(-0.5, -0.5 * a, 1.0) @ trans(-1.0, -1.0 / a, 0.0) @ perspective(fovy, 1.0, 1e-2, 1000.0) @ scale(1.0 if flip else -1.0, 1.0, -1.0)
        expect = np.linalg.inv(c2s)
        got = ob.camera_new(w, h, fov, np.eye(4, dtype=np.float32), fov_axis=axis, flip=flip).reshape(4, 4).T.astype(np.float64)  # column-major
        scale_ = np.abs(expect).max()
        assert np.allclose(got, expect, rtol=2e-5, atol=2e-6 * scale_), (w, h, fov, axis, flip, np.abs(got - expect).max())
