// mxb_ops.cuh — element physics and array search of the MARXS hot path, fp64, sm_100a.
//
// These are the op bodies.  Two kernels are built from them:
//   * the interpreter (mxb_trace.cu: mxb_trace_kernel), which dispatches on the op list of a
//     program blob at run time, and
//   * the per-program specialised kernels (mxb_jit.cpp), whose straight-line driver is emitted
//     from the same op list and compiled with NVRTC for sm_100a: offsets, flags, column
//     presence and draw sources are compile-time constants there and single-element
//     parameters are read from the kernel-parameter constant bank instead of shared memory.
// Every body restates one reference routine (file:line cited), in the operation order of
// oracle/marxs_oracle.py.  Random draws are ARGUMENTS: the caller decides between injected
// per-photon arrays and the device Philox stream.
#pragma once
#ifdef __CUDACC_RTC__
#include "mxb.h"
#else
#include "../../include/mxb.h"
#endif
#include "mxb_device.cuh"

namespace mxb {

// ---------------------------------------------------------------------------
// TMA bulk copy global -> shared (SASS: UBLKCP), completion on an mbarrier
// ---------------------------------------------------------------------------
MXB_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
MXB_DEV void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
MXB_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
MXB_DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
MXB_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

// ---------------------------------------------------------------------------
// parameter access.  PRef: word offsets into the staged program copy (shared memory, 32-bit
// offsets -> LDS) or into the global blob when the program is too large to stage.
// SRef: a block of kernel parameters (constant bank; offsets fold into the instruction).
// ---------------------------------------------------------------------------
extern __shared__ __align__(16) double g_smem[];

template <bool STAGED>
struct PRef {
    const double* g;
    int off;
    MXB_DEV double operator[](int k) const {
        if (STAGED) return g_smem[off + k];
        return __ldg(g + off + k);
    }
    MXB_DEV PRef operator+(int k) const { return PRef{g, off + k}; }
    MXB_DEV double2 ld2(int k) const {  // words 2k, 2k+1 (off must be even)
        if (STAGED) return reinterpret_cast<const double2*>(g_smem + off)[k];
        return __ldg(reinterpret_cast<const double2*>(g + off) + k);
    }
    MXB_DEV int i32(int k) const {  // packed int32 view of the words at off
        if (STAGED) return reinterpret_cast<const int*>(g_smem)[2 * off + k];
        return __ldg(reinterpret_cast<const int*>(g) + 2 * (long long)off + k);
    }
};

struct SRef {
    const double* s;
    MXB_DEV double operator[](int k) const { return s[k]; }
    MXB_DEV SRef operator+(int k) const { return SRef{s + k}; }
    MXB_DEV double2 ld2(int k) const { return make_double2(s[2 * k], s[2 * k + 1]); }
};

MXB_DEV double nan64() { return __longlong_as_double(0x7ff8000000000000LL); }

// ---------------------------------------------------------------------------
// per-thread photon state
// ---------------------------------------------------------------------------
struct Photon {
    V3 pos, dir, pol;
    double energy, prob;
    V3 ip;          // intersection point of the current element
    double l0, l1;  // local coordinates on the current element
    double last_order;   // order drawn by the latest GRATING op (QualityFactor reads photons['order'])
    bool hit;
    bool unit;      // |dir| == 1 to rounding (fast build: see kTrackUnit in mxb_device.cuh)
};
MXB_DEV void photon_loaded(Photon& ph) {   // after pos/dir/pol/energy/prob are in registers
    ph.hit = false;
    ph.last_order = nan64();
    ph.unit = kTrackUnit && (dot(ph.dir, ph.dir) == 1.0);
}

// ---------------------------------------------------------------------------
// per-warp input pipeline.  Each warp owns 11 x 32 doubles of shared memory and one mbarrier; lane 0
// fetches the NEXT group of 32 photons (11 core planes, 256 B each) with TMA bulk copies while the
// warp traces the current one, so a warp never waits on DRAM latency at the top of its loop and no
// registers are spent on double buffering.  Needs 16-byte aligned planes and a full group; anything
// else (tail, odd plane offsets) is read with ordinary loads.
// ---------------------------------------------------------------------------
#define MXB_IN_PLANES 11
#define MXB_PIPE_WORDS_PER_WARP (MXB_IN_PLANES * 32)
struct InputPipe {
    double* buf;      // this warp's slice
    uint64_t* bar;    // this warp's mbarrier (count 1: lane 0's expect_tx arrive)
    unsigned phase;
    bool pending;     // a group is in flight / has landed in buf
};
MXB_DEV void pipe_issue(InputPipe& p, const double* const* planes, long long base) {   // lane 0 only
    mbar_expect_tx(p.bar, MXB_IN_PLANES * 256u);
#pragma unroll
    for (int k = 0; k < MXB_IN_PLANES; ++k) bulk_g2s(p.buf + 32 * k, planes[k] + base, 256u, p.bar);
}
MXB_DEV void pipe_start(InputPipe& p, const double* const* planes, long long base, long long n, bool tma_ok, int lane) {
    p.phase = 0u;
    p.pending = tma_ok && (base + 32 <= n);
    if (p.pending && lane == 0) pipe_issue(p, planes, base);
}
// photon `base + lane` into registers, then start fetching the group at next_base.  Lanes past the
// end of the batch re-read the last photon (their results are never stored).
MXB_DEV void pipe_load(InputPipe& p, const double* const* planes, long long base, long long next_base, long long n,
                       bool tma_ok, int lane, bool active, V3& pos, V3& dir, V3& pol, double& energy, double& prob) {
    if (p.pending) {
        mbar_wait(p.bar, p.phase);
        p.phase ^= 1u;
        const double* b = p.buf + lane;
        pos = V3{b[0], b[32], b[64]};
        dir = V3{b[96], b[128], b[160]};
        pol = V3{b[192], b[224], b[256]};
        energy = b[288];
        prob = b[320];
        __syncwarp();   // every lane has read its values before the slice is refilled
    } else {
        const long long il = active ? (base + lane) : (n - 1);
        pos = V3{planes[0][il], planes[1][il], planes[2][il]};
        dir = V3{planes[3][il], planes[4][il], planes[5][il]};
        pol = V3{planes[6][il], planes[7][il], planes[8][il]};
        energy = planes[9][il];
        prob = planes[10][il];
    }
    p.pending = tma_ok && (next_base + 32 <= n);
    if (p.pending && lane == 0) pipe_issue(p, planes, next_base);
}


MXB_DEV void st_global(double* p, double v) {
    asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
MXB_DEV void st_global(long long* p, long long v) {
    asm volatile("st.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// status counters are bumped on rare paths from many sites: one shared routine
// (every call site passes a constant, so the lanes that arrive here together count the same word: one
//  shared-memory atomic per converged group instead of one per lane - steep-ray photons of a strongly
//  dispersing array would otherwise serialise on one address)
// (out of line: it is called from ~20 cold sites of a fused program, and inlined there it sits between the hot blocks;
//  measured on C2: 0.891 -> 0.879 ms)
#ifndef MXB_INLINE_STATUS
__device__ __noinline__
#else
MXB_DEV
#endif
void count_status(unsigned long long* st_sm, int which) {
    const unsigned m = __activemask();
    if ((int)(threadIdx.x & 31u) == __ffs(m) - 1) atomicAdd(&st_sm[which], (unsigned long long)__popc(m));
}

// optics/base.py:43-47: probability factors multiply and must lie in [0,1]
MXB_DEV void mul_prob(unsigned long long* st_sm, Photon& ph, double f) {
    if (f < 0.0 || f > 1.0) count_status(st_sm, MXB_ST_PROB_RANGE);
    ph.prob *= f;
}

// draw source: injected per-photon array (tests, parity) or Philox keyed by (seed, photon id, slot)
MXB_DEV double draw_value(const double* inj, long long i, unsigned long long seed, unsigned long long gid,
                          int slot, int kind) {
    if (inj) return inj[i];
    return device_draw(seed, gid, slot, kind);
}

// ---------------------------------------------------------------------------
// detector image accumulation.  A bright spot (zero order, a narrow line) sends a large fraction of
// the batch into a handful of pixels; fp64 atomics on ONE address serialise in L2 (~1 ns each), which
// costs more than the whole trace.  So: (1) lanes of a warp that hit the same pixel are summed with
// shuffles and issue one atomic; (2) each CTA keeps a small direct-mapped cache of hot pixels in
// shared memory - a pixel is admitted when two lanes of one warp coincide on it - and flushes it
// once at the end of the kernel.  Cold pixels go straight to global memory.
// ---------------------------------------------------------------------------
#define MXB_HOT_SLOTS 512
static_assert(MXB_HOT_SLOTS == 512, "hot_add takes the top 9 bits of the pixel hash as the slot");
struct HotCache {
    unsigned long long* keys;   // global address of the pixel, 0 = free
    double* vals;
};
MXB_DEV void hot_init(HotCache hc, int tid, int nthreads) {
    for (int k = tid; k < MXB_HOT_SLOTS; k += nthreads) {
        hc.keys[k] = 0ULL;
        hc.vals[k] = 0.0;
    }
}
MXB_DEV void hot_flush(HotCache hc, int tid, int nthreads) {   // after a __syncthreads()
    for (int k = tid; k < MXB_HOT_SLOTS; k += nthreads)
        if (hc.keys[k]) atomicAdd(reinterpret_cast<double*>(hc.keys[k]), hc.vals[k]);
}
// pix = linear pixel index of addr in its image (the 32-bit key the lanes are matched on)
MXB_DEV void hot_add(HotCache hc, double* addr, double w, unsigned pix) {
    const unsigned long long key = (unsigned long long)addr;
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, pix);
    const int leader = __ffs(peers) - 1;
    double sum = w;
    unsigned rest = peers & ~(1u << leader);
    while (rest) {   // every lane of the group walks the same list
        const int l = __ffs(rest) - 1;
        sum += __shfl_sync(peers, w, l);
        rest &= rest - 1;
    }
    if ((int)(threadIdx.x & 31) != leader) return;
    const unsigned slot = (pix * 0x9E3779B9u) >> 23;     // MXB_HOT_SLOTS = 512
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&hc.keys[slot]);
    if (cur == 0ULL && (peers & (peers - 1))) {   // free slot and the pixel looks hot: claim it
        cur = atomicCAS(&hc.keys[slot], 0ULL, key);
        if (cur == 0ULL) cur = key;
    }
    if (cur == key) atomicAdd(&hc.vals[slot], sum);
    else atomicAdd(addr, sum);
}

// fused detector image: gp = nx ny sel_lo n_sel ; bin = round half to even like np.round
template <typename PP>
MXB_DEV void accumulate_image(HotCache hc, double* img, PP gp, long long idn, double px, double py, double w) {
    const long long nx = (long long)gp[0], ny = (long long)gp[1];
    const long long plane = idn - (long long)gp[2];
    if (plane < 0 || plane >= (long long)gp[3] || !(px == px) || !(py == py) || !(w == w)) return;
    const long long ix = llrint(px), iy = llrint(py);
    if (ix < 0 || iy < 0 || ix >= nx || iy >= ny) return;
    const long long lin = (plane * ny + iy) * nx + ix;
    hot_add(hc, &img[lin], w, (unsigned)lin);
}

// ---------------------------------------------------------------------------
// element physics (callers only invoke these for photons with ph.hit)
// ---------------------------------------------------------------------------
// mirror.py:53-82  params: P[3] f
template <typename PP>
MXB_DEV void op_lens(Photon& ph, PP p) {
    const V3 nd = normalize_unless(ph.unit, ph.dir);
    const double f = p[3];
    const V3 t{(p[0] + f * nd.x) - ph.ip.x, (p[1] + f * nd.y) - ph.ip.y, (p[2] + f * nd.z) - ph.ip.z};
    const V3 nd2 = normalize(t);
    ph.pol = parallel_transport(kTrackUnit ? nd : ph.dir, nd2, ph.pol, kTrackUnit, true);
    ph.dir = nd2;
    ph.unit = true;
}

// mirror.py:68-81  PerfectLens with a reflectivity_interpolator: probability *= R(energy, angle / 4)^2,
// angle = arccos|d_new . d_old|, R = RectBivariateSpline(kx=ky=1).ev = bilinear with the query clamped to
// the table.  params: P[3] f nx ny table ; table (global) = x[nx] y[ny] z[nx][ny]
template <typename PP>
MXB_DEV void op_lens_refl(unsigned long long* st_sm, Photon& ph, PP p, const double* gprog) {
    const V3 nd = normalize_unless(ph.unit, ph.dir);
    op_lens(ph, p);
    const double angle = m_acos(fabs(dot(ph.dir, nd)));
    const int nx = (int)p[4], ny = (int)p[5];
    const PRef<false> xk{gprog, (int)p[6]};
    const PRef<false> yk = xk + nx;
    const PRef<false> z = yk + ny;
    const double xq = fmin(fmax(ph.energy, xk[0]), xk[nx - 1]);
    const double yq = fmin(fmax(angle / 4, yk[0]), yk[ny - 1]);
    const int i = bracket(xk, nx, xq), j = bracket(yk, ny, yq);
    const double tx = div(xq - xk[i], xk[i + 1] - xk[i]);
    const double ty = div(yq - yk[j], yk[j + 1] - yk[j]);
    const double a00 = z[i * ny + j], a10 = z[(i + 1) * ny + j], a01 = z[i * ny + j + 1], a11 = z[(i + 1) * ny + j + 1];
    const double f0 = a00 + tx * (a10 - a00);
    const double f1 = a01 + tx * (a11 - a01);
    const double r = f0 + ty * (f1 - f0);
    mul_prob(st_sm, ph, r * r);
}

// scatter.py:49-77  params: center[3] sig_in sig_perp ; z0, z1 standard normal draws
// NZ_IN / NZ_PERP: whether the width is non-zero, when the caller knows at compile time (1 / 0; -1: test p[3] / p[4])
template <int NZ_IN = -1, int NZ_PERP = -1, typename PP>
MXB_DEV void op_rscatter(Photon& ph, PP p, double z0, double z1, double& a, double& b) {
    const V3 radial{ph.pos.x - p[0], ph.pos.y - p[1], ph.pos.z - p[2]};
    const V3 perp = cross(ph.dir, radial);
    V3 out = ph.dir;
    a = 0.0;
    b = 0.0;
    if (NZ_IN < 0 ? (p[3] != 0.0) : (NZ_IN != 0)) {
        a = p[3] * z0;
        out = axangle_rotate_T<true>(perp, a, ph.dir);     // perp = dir x radial
    }
    if (NZ_PERP < 0 ? (p[4] != 0.0) : (NZ_PERP != 0)) {
        b = p[4] * z1;
        out = axangle_rotate_T(radial, b, out);
    }
    ph.pol = parallel_transport(ph.dir, out, ph.pol, ph.unit, ph.unit);   // rotations keep |dir|
    ph.dir = out;
}
// the two normals of op_rscatter: drawn only for non-zero widths; one Philox call when both come
// from the device stream
template <int NZ_IN = -1, int NZ_PERP = -1>
MXB_DEV void rscatter_draws(double sig_in, double sig_perp, const double* inj0, const double* inj1, long long i,
                            unsigned long long seed, unsigned long long gid, int s0, int s1, double& z0,
                            double& z1) {
    z0 = 0.0;
    z1 = 0.0;
    const bool nz_in = NZ_IN < 0 ? (sig_in != 0.0) : (NZ_IN != 0);
    const bool nz_perp = NZ_PERP < 0 ? (sig_perp != 0.0) : (NZ_PERP != 0);
    if (nz_in && nz_perp && !inj0 && !inj1) {
        device_draw_normal_pair(seed, gid, s0, z0, z1);
    } else {
        if (nz_in) z0 = draw_value(inj0, i, seed, gid, s0, 1);
        if (nz_perp) z1 = draw_value(inj1, i, seed, gid, s1, 1);
    }
}

// scatter.py:109-145  params: sigma ; zn standard normal, u uniform
template <typename PP>
MXB_DEV void op_gscatter(Photon& ph, PP p, int flags, double zn, double u, double& ang) {
    const V3 pdir = normalize_unless(ph.unit, ph.dir);
    const V3 guess = (fabs(pdir.x) < 0.99999) ? V3{1, 0, 0} : V3{0, 1, 0};
    const V3 perp = cross(pdir, guess);
    if (flags & 2) {   // callable scatter (scatter.py:127-129): the caller passes the angle of this photon in zn
        ang = zn;
    } else if (flags & 1) {   // L2Diffraction (mitsnl/catgrating.py:280-285): Airy-disk sigma of the L2 mesh, p[0] = innerfree
        const double wave = div(kHcKevNm * 1e-6, ph.energy);   // astropy u.spectral(): keV -> mm
        const double sigma = (1.22 * 0.4) * m_asin(div(wave, p[0]));
        ang = zn * sigma;
    } else {
        ang = p[0] * zn;
    }
    V3 out = axangle_rotate_T<true>(perp, ang, pdir);   // perp = pdir x guess
    double s2, c2;
    sincos_turn(u, &s2, &c2);                           // ang2 = u * 2 * pi
    out = rotate_T_sc<false, true>(pdir, s2, c2, out);  // pdir was normalised above
    ph.pol = parallel_transport(kTrackUnit ? pdir : ph.dir, out, ph.pol, kTrackUnit, true);
    ph.dir = out;
    ph.unit = true;
}

// filter.py:90-94  params: n, x[n], y[n] [, fill_lo, fill_hi]   (n == 0: constant y[0])
// flags (scipy.interpolate.interp1d modes): 1 bounds_error -> ValueError outside the table; 2 out-of-range
// queries return fill_lo / fill_hi (fill_value=number or pair; NaN by default); 4 fill_value='extrapolate'
// (the end segments continue); none of them: ends clamped like np.interp
template <typename PP>
MXB_DEV double filter_value(unsigned long long* st_sm, PP p, double energy, int flags) {
    const int n = (int)p[0];
    if (n == 0) return p[1];
    PP xp = p + 1;
    PP fp = p + 1 + n;
    const bool below = energy < xp[0], above = energy > xp[n - 1];
    if (below || above) {
        if (flags & 1) count_status(st_sm, MXB_ST_FILTER_BOUNDS);
        else if (flags & 2) return below ? p[1 + 2 * n] : p[2 + 2 * n];
        else if ((flags & 4) && n >= 2) {
            const int lo = below ? 0 : n - 2;
            const double slope = div(fp[lo + 1] - fp[lo], xp[lo + 1] - xp[lo]);
            return slope * (energy - xp[lo]) + fp[lo];
        }
    }
    return interp_clamped(xp, fp, n, energy);
}

// grating.py:12-57 OrderSelector with a compile-time order count (specialised kernels):
// sel = kind n psum cdf[n] orders[n]
template <int N, typename PP>
MXB_DEV double select_order_fixed(PP sel, double u, double& psel) {
    psel = sel[2];
    int idx = 0;   // searchsorted(cdf, u, 'right'), clamped
#pragma unroll
    for (int k = 0; k < N - 1; ++k) idx += (sel[3 + k] <= u) ? 1 : 0;
    double order = sel[3 + N];
#pragma unroll
    for (int k = 1; k < N; ++k) order = (idx == k) ? sel[3 + N + k] : order;
    return order;
}

// grating.py:12-57, 60-96; mitsnl/catgrating.py:104-144.  Returns order, sets psel.
template <typename PP>
MXB_DEV double select_order(PP sel, const double* gprog, double u, double energy, double blaze, double& psel) {
    const int kind = (int)sel[0];
    if (kind == MXB_SEL_ORDERSELECTOR) {
        const int n = (int)sel[1];
        psel = sel[2];
        PP cdf = sel + 3;
        int idx = 0;   // searchsorted(cdf, u, 'right') = number of entries <= u (clamped), branch-free
        for (int k = 0; k < n - 1; ++k) idx += (cdf[k] <= u) ? 1 : 0;
        return sel[3 + n + idx];
    } else if (kind == MXB_SEL_EFFFILE) {
        const int nE = (int)sel[1], nO = (int)sel[2];
        PP en = sel + 3;
        int ind = 0;
        double best = fabs(en[0] - energy);
        for (int k = 1; k < nE; ++k) {  // np.argmin: first minimum
            const double d = fabs(en[k] - energy);
            if (d < best) { best = d; ind = k; }
        }
        psel = sel[3 + nE + ind];
        PP cum = sel + 3 + 2 * nE + nO + ind * nO;
        int oi = 0;
        for (int k = 0; k < nO; ++k)
            if (cum[k] > u) { oi = k; break; }
        return sel[3 + 2 * nE + oi];
    } else {  // MXB_SEL_INTERPTABLE: bilinear, query clamped to the table (RectBivariateSpline k=1)
        const int nw = (int)sel[1], nt = (int)sel[2], no = (int)sel[3];
        const double* tab = gprog + (long long)sel[4];
        PP wk = sel + 5;
        PP tk = sel + 5 + nw;
        PP ord = sel + 5 + nw + nt;
        double xq = div(kHcKevNm, energy);
        xq = fmin(fmax(xq, wk[0]), wk[nw - 1]);
        double yq = fmin(fmax(blaze, tk[0]), tk[nt - 1]);
        const int i = bracket(wk, nw, xq), j = bracket(tk, nt, yq);
        const double tx = div(xq - wk[i], wk[i + 1] - wk[i]);
        const double ty = div(yq - tk[j], tk[j + 1] - tk[j]);
        const double* t00 = tab + ((long long)i * nt + j) * no;
        const double* t10 = t00 + (long long)nt * no;
        const double* t01 = t00 + no;
        const double* t11 = t10 + no;
#ifdef MXB_FAST
        // Fast build: interpolation is linear, so the interpolated running sum over orders equals the
        // running sum of the interpolated efficiencies up to rounding.  With the cumulative table the
        // total is ONE bilinear lookup and the order a bisection over k (5 lookups for 28 orders)
        // instead of 2 x 28 lookups; the order can only differ from the reference when u sits within
        // ~1e-16 of a bin boundary.
        {
            const double* c00 = t00 + (long long)nw * nt * no;
            const double* c10 = c00 + (long long)nt * no;
            const double* c01 = c00 + no;
            const double* c11 = c10 + no;
            auto cum = [&](int k) {
                const double a00 = __ldg(c00 + k), a10 = __ldg(c10 + k), a01 = __ldg(c01 + k), a11 = __ldg(c11 + k);
                const double f0 = a00 + tx * (a10 - a00);
                const double f1 = a01 + tx * (a11 - a01);
                return f0 + ty * (f1 - f0);
            };
            const double total = cum(no - 1);
            psel = total;
            int lo = 0, hi = no - 1;
            if (!(total / total > u)) {
                lo = 0;                         // no cumulative fraction exceeds u: np.argmax of all-False is 0
            } else {
                const double rt = fast_rcp(total);
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (cum(mid) * rt > u) hi = mid; else lo = mid + 1;
                }
            }
            return ord[lo];
        }
#endif
        // first pass: interpolated efficiency of every order (kept in local memory for the second pass
        // when there are at most 32 orders) and their sum; second pass: first order whose cumulative
        // fraction exceeds u
        constexpr int kMaxCached = 32;
        double fv[kMaxCached];
        const bool cached = no <= kMaxCached;
        double total = 0.0;
        for (int k = 0; k < no; ++k) {
            const double a00 = __ldg(t00 + k), a10 = __ldg(t10 + k), a01 = __ldg(t01 + k), a11 = __ldg(t11 + k);
            const double f0 = a00 + tx * (a10 - a00);
            const double f1 = a01 + tx * (a11 - a01);
            const double f = f0 + ty * (f1 - f0);
            if (cached) fv[k] = f;
            total = total + f;
        }
        psel = total;
        double run = 0.0;
        int oi = 0;
        for (int k = 0; k < no; ++k) {
            double f;
            if (cached) {
                f = fv[k];
            } else {
                const double a00 = __ldg(t00 + k), a10 = __ldg(t10 + k), a01 = __ldg(t01 + k), a11 = __ldg(t11 + k);
                const double f0 = a00 + tx * (a10 - a00);
                const double f1 = a01 + tx * (a11 - a01);
                f = f0 + ty * (f1 - f0);
            }
            run = run + f;
            if (run / total > u) { oi = k; break; }   // argmax(cumprob > u): first True, 0 if none
        }
        return ord[oi];
    }
}

// grating.py:233-277  params: l[3] dd[3] d blaze0 dblaze ; n = e_x of the geometry.
// flags: 1 CAT sign convention, 2 reflection, 4 blaze modifier, 8 L1 support bars, 16 per-photon d (column)
// select(energy, blaze, psel) -> diffraction order (the caller binds the draw and the table)
template <typename PP, typename GP, typename SELECT>
MXB_DEV void op_grating(unsigned long long* st_sm, Photon& ph, PP p, GP geom, int flags, SELECT select,
                        double& order, double& blaze, bool l1_blocked = false, double l1_trans = 0.0,
                        double d_photon = 0.0) {
    const V3 pn = normalize_unless(ph.unit, ph.dir);
    const V3 l = ld3(p), dd = ld3(p + 3), n = ld3(geom + 3);
    const double wave = div(kEnergy2Wave, ph.energy);
    const double p_l = dot(pn, l);
    const V3 pp = normalize(V3{pn.x - p_l * l.x, pn.y - p_l * l.y, pn.z - p_l * l.z});
    blaze = m_acos(clip01(fabs(dot(pp, n))));
    if (flags & 4) blaze = blaze + (p[7] + ph.l0 * p[8]);  // NonParallelCATGrating blaze_angle_modifier
    double psel;
    order = select(ph.energy, blaze, psel);
    const double p_dd = dot(pn, dd);
    const double sign = (flags & 1) ? ((p_dd < 0.0) ? -1.0 : 1.0) : -1.0;  // CAT: grating.py:298-301
    // flags bit 4: the grating constant varies over the facet (grating.py:209-220, callable d): the caller
    // passes the value d(intercoos) of this photon
    const double p_d = p_dd + div(sign * order * wave, (flags & 16) ? d_photon : p[6]);
    const double p_n = m_sqrt(1. - p_d * p_d - p_l * p_l);     // NaN for an evanescent order, like np.sqrt
    const double pdn = dot(pn, n);
    double direction = (pdn > 0.0) ? 1.0 : ((pdn < 0.0) ? -1.0 : pdn);  // np.sign
    if (flags & 2) direction = direction * -1;
    const double q = direction * p_n;
    const V3 nd{p_d * dd.x + p_l * l.x + q * n.x, p_d * dd.y + p_l * l.y + q * n.y,
                p_d * dd.z + p_l * l.z + q * n.z};
    if (l1_blocked) {
        // L1 support bar (mitsnl/catgrating.py:203-218): the photon goes through solid Si instead of the
        // open grating: direction and polarization stay, order 0, probability = Si transmission
        order = 0.0;
        ph.last_order = order;
        mul_prob(st_sm, ph, l1_trans);
        return;
    }
    ph.last_order = order;
    // nd is a unit vector by construction (p_d^2 + p_l^2 + p_n^2 = 1 in the orthonormal frame d, l, n)
    ph.pol = parallel_transport(kTrackUnit ? pn : ph.dir, nd, ph.pol, kTrackUnit, true);
    ph.dir = nd;
    ph.unit = true;
    mul_prob(st_sm, ph, psel);
}

// mitsnl/catgrating.py:147-161  params: factor
#ifdef MXB_FAST
__device__ __noinline__ double pow_cold(double x, double y) { return m_pow(x, y); }
#endif
// factor ** y with log(factor) = lhi + llo supplied by the host as a double-double (the factor is one number per
// element: missions/mitsnl/catgrating.py lowers it with 40-digit arithmetic).  Fast build: exp(y lhi) corrected by the
// rounding of the product and by y llo, i.e. libm's exp (1 ulp) instead of libm's pow (~200 instructions with an
// extended-precision log inside); <= 3 ulp for any exponent.  A NaN lhi (factor <= 0 or not finite) takes libm's pow.
MXB_DEV double pow_loghost(double x, double lhi, double llo, double y) {
#ifdef MXB_FAST
    if (lhi == lhi) {
        if (lhi == 0.0 && llo == 0.0) return 1.0;        // 1 ** y
        const double p = y * lhi;
        const double e = fma(y, lhi, -p) + y * llo;
        const double r = m_exp(p);
        return fma(r, e, r);
    }
    return pow_cold(x, y);
#else
    return m_pow(x, y);
#endif
}
// mitsnl/catgrating.py:147-161  params: factor, log(factor) hi, lo
template <typename PP>
MXB_DEV void op_qfactor(unsigned long long* st_sm, Photon& ph, PP p) {
    mul_prob(st_sm, ph, pow_loghost(p[0], p[1], p[2], ph.last_order * ph.last_order));
}

// mitsnl/catgrating.py:222-259  params: openfraction, bardepth * innerfree, totalarea ; n = e_x of the geometry
template <typename PP, typename GP>
MXB_DEV void op_l2abs(unsigned long long* st_sm, Photon& ph, PP p, GP geom) {
    const V3 p3 = normalize_unless(ph.unit, ph.dir);
    const V3 en = ld3(geom + 3);
#ifdef MXB_OWN_ASIN
    // sin(arccos(c)) = sqrt((1 - c)(1 + c)) (1 - c is exact for c >= 1/2); NaN above 1 like the reference
    const double c = fabs(dot(p3, en));
    const double sn = sqrt_nn((1.0 - c) * (1.0 + c));
#else
    const double angle = m_acos(fabs(dot(p3, en)));   // no clip in the reference: NaN above 1
    const double sn = m_sin(angle);
#endif
    mul_prob(st_sm, ph, p[0] - div(p[1] * sn, p[2]));
}

// multiLayerMirror.py:44-91  params: Pinv[9] P[9] ex[3]
template <typename PP>
MXB_DEV void op_brewster(unsigned long long* st_sm, Photon& ph, PP p) {
    const V3 dh = normalize(ph.dir);
    V3 loc{p[0] * dh.x + p[1] * dh.y + p[2] * dh.z, p[3] * dh.x + p[4] * dh.y + p[5] * dh.z,
           p[6] * dh.x + p[7] * dh.y + p[8] * dh.z};
    loc.x = loc.x * -1;
    PP q = p + 9;
    const V3 nd{q[0] * loc.x + q[1] * loc.y + q[2] * loc.z, q[3] * loc.x + q[4] * loc.y + q[5] * loc.z,
                q[6] * loc.x + q[7] * loc.y + q[8] * loc.z};
    const V3 ex = ld3(p + 18);
    const V3 v_s = normalize(cross(dh, ex));
    const V3 v_p = cross(dh, v_s);
    const double pvs = dot(ph.pol, v_s), pvp = dot(ph.pol, v_p);
    const double Es2 = 1. * (pvs * pvs), Ep2 = 0. * (pvp * pvp);
    const double inten = Es2 + Ep2;
    if (inten > 1.001) count_status(st_sm, MXB_ST_INTENSITY);
    const V3 nvp = cross(nd, v_s);
    const V3 np_{-Es2 * v_s.x + Ep2 * nvp.x, -Es2 * v_s.y + Ep2 * nvp.y, -Es2 * v_s.z + Ep2 * nvp.z};
    ph.pol = normalize(np_);
    ph.dir = nd;
    ph.unit = false;   // pos4d may carry zoom / shear: renormalise at the next use
    mul_prob(st_sm, ph, clip01(inten));
}

// multiLayerMirror.py:132-170  params: Ly n_refl n_pol xs[nr] peak_lambda[nr] peak[nr] fwhm[nr] pol_e[np] pol[np]
template <typename PP>
MXB_DEV void op_mleff(unsigned long long* st_sm, Photon& ph, PP p) {
    const double Ly = p[0];
    const int nr = (int)p[1], npol = (int)p[2];
    PP xs = p + 3;
    PP pl = xs + nr;
    PP pk = pl + nr;
    PP fw = pk + nr;
    PP pe = fw + nr;
    PP pf = pe + npol;
    const double wavelength = div(kHcMultilayer, ph.energy);
    const double tested = interp_clamped(pe, pf, npol, ph.energy);
    const double local_x = div(ph.l0, Ly);
    // three columns at one position: one bisection (the values are interp_clamped's, bit for bit)
    int lo;
    const int mode = interp_bracket(xs, nr, local_x, lo);
    const double x0 = mode ? 0.0 : xs[lo], dx = mode ? 1.0 : xs[lo + 1] - x0;      // (unused when clamped)
    const double peak_w = interp_on_bracket(pl, nr, mode, lo, local_x, x0, dx);
    const double max_refl = div(interp_on_bracket(pk, nr, mode, lo, local_x, x0, dx), tested);
    const double spread = interp_on_bracket(fw, nr, mode, lo, local_x, x0, dx);
    const double c2 = div(spread * spread, 8. * 0.6931471805599453);
    double refl = 0.0;
    if (c2 != 0.0) {
        const double dw = wavelength - peak_w;
        refl = max_refl * m_exp(div(-(dw * dw), 2 * c2));
    }
    mul_prob(st_sm, ph, div(refl, 100));
}

// detector.py:73-75  pr: pixsize cp0 cp1
template <typename PP>
MXB_DEV void op_detpix(const Photon& ph, PP pr, int flags, double& px, double& py) {
    double qx, qy;
    div2((flags & 2) ? ph.l0 * pr[3] : ph.l0, ph.l1, pr[0], qx, qy);   // CircularDetector: phi * R / pixsize (detector.py:114)
    px = qx + pr[1];
    py = qy + pr[2];
}

// ---------------------------------------------------------------------------
// math/geometry.py:470-564 Cylinder.intersect: unit circle in the local xy plane, |z| <= 1.
// g: inv(pos4d)[16] pos4d[16] (row major) phi_lo phi_hi zoom_z.  Of the two roots the nearer valid
// one (a >= 0, |z| <= 1, phi inside the limits) wins; local coordinates are (phi, z * zoom_z).
// ---------------------------------------------------------------------------
MXB_DEV double py_mod(double a, double m) {   // np.remainder: result carries the sign of m
    double r = fmod(a, m);
    if (r != 0.0 && ((r < 0.0) != (m < 0.0))) r += m;
    return r;
}
MXB_DEV bool angle_between(double angle, double b1, double b2) {   // math/utils.py:180-215, borders pre-normalised
    const double ang = py_mod(kTwoPi + py_mod(angle, kTwoPi), kTwoPi);
    return (b1 < b2) ? ((b1 <= ang) && (ang <= b2)) : ((b1 <= ang) || (ang <= b2));
}
template <typename P>
MXB_DEV bool cylinder_intersect(P g, const V3& pos, const V3& dir, V3& ip, double& l0, double& l1) {
    double dl[3], pl[4];
#pragma unroll
    for (int r = 0; r < 3; ++r) dl[r] = g[4 * r] * dir.x + g[4 * r + 1] * dir.y + g[4 * r + 2] * dir.z;
#pragma unroll
    for (int r = 0; r < 4; ++r) pl[r] = g[4 * r] * pos.x + g[4 * r + 1] * pos.y + g[4 * r + 2] * pos.z + g[4 * r + 3];
    const double x = div(pl[0], pl[3]), y = div(pl[1], pl[3]), z = div(pl[2], pl[3]);
    const double c = (x * x + y * y) - 1.;
    const double b = 2 * (x * dl[0] + y * dl[1]);
    const double a = dl[0] * dl[0] + dl[1] * dl[1];
    const double underroot = b * b - 4 * a * c;
    const bool real = underroot >= 0;
    const double sq = m_sqrt(underroot);
    const double denom = 2 * a;
    const double a1 = div(-b + sq, denom), a2 = div(-b - sq, denom);
    const double x1 = x + a1 * dl[0], y1 = y + a1 * dl[1], z1 = z + a1 * dl[2];
    const double x2 = x + a2 * dl[0], y2 = y + a2 * dl[1], z2 = z + a2 * dl[2];
    const double phi1 = m_atan2(y1, x1), phi2 = m_atan2(y2, x2);
    bool hit1 = real && (a1 >= 0) && (fabs(z1) <= 1.) && angle_between(phi1, g[32], g[33]);
    bool hit2 = real && (a2 >= 0) && (fabs(z2) <= 1.) && angle_between(phi2, g[32], g[33]);
    hit1 = hit1 && !(hit2 && (a2 < a1));   // both valid: the closer one
    hit2 = hit2 && !(hit1 && (a2 >= a1));
    const double lx = hit1 ? x1 : x2, ly = hit1 ? y1 : y2, lz = hit1 ? z1 : z2;
    l0 = hit1 ? phi1 : phi2;
    l1 = lz * g[34];
    P q = g + 16;
    ip.x = q[0] * lx + q[1] * ly + q[2] * lz + q[3];
    ip.y = q[4] * lx + q[5] * ly + q[6] * lz + q[7];
    ip.z = q[8] * lx + q[9] * ly + q[10] * lz + q[11];
    return hit1 || hit2;
}

// det_acis.py:31-58 + data.py:169-190 ; per-facet pr: pixsize cp0 cp1 sh ct st ox oy ;
// global gp: f pixrad odet0 odet1 cosr sinr.  out: chipx chipy tdetx tdety detx dety x y
template <typename PP, typename GP>
MXB_DEV void op_acis(const Photon& ph, PP pr, GP gp, double out[8]) {
    double cx, cy, xm, ym, x, y;
    div2(ph.l0, ph.l1, pr[0], cx, cy);
    const double chipx = cx + pr[1] + 1;
    const double chipy = cy + pr[2] + 1;
    const double tx = pr[3] * (pr[4] * (chipx - 0.5) + pr[5] * (chipy - 0.5)) + pr[6];
    const double ty = pr[3] * (-pr[5] * (chipx - 0.5) + pr[4] * (chipy - 0.5)) + pr[7];
    const double mn0 = ph.ip.x - gp[0];
    div2(ph.ip.y, ph.ip.z, mn0, xm, ym);
    div2(xm, ym, gp[1], x, y);
    out[0] = chipx;
    out[1] = chipy;
    out[2] = tx;
    out[3] = ty;
    out[4] = gp[2] - x;
    out[5] = gp[3] + y;
    out[6] = gp[2] - x * gp[4] + y * gp[5];
    out[7] = gp[3] + x * gp[5] + y * gp[4];
}

// aperture.py:42-78: params c[3] vy[3] vz[3] nex[3] phi0 dphi rin2 cum_lo cum_hi ; u0, u1 uniforms
template <typename PP>
MXB_DEV void op_aperture(unsigned long long* st_sm, Photon& ph, PP pr, int flags, double u0, double u1) {
    double x, y;
    if (flags & 1) {  // CircleAperture :138-146
        const double phi = pr[12] + pr[13] * u0;
        const double r = m_sqrt(pr[14] + (1. - pr[14]) * u1);
        double sn, cs;
        m_sincos(phi, &sn, &cs);
        x = r * cs;
        y = r * sn;
    } else {  // RectangleAperture :92-95
        x = u0 * 2. - 1.;
        y = u1 * 2. - 1.;
    }
    ph.l0 = x;
    ph.l1 = y;
    ph.ip = V3{pr[0] + x * pr[3] + y * pr[6], pr[1] + x * pr[4] + y * pr[7], pr[2] + x * pr[5] + y * pr[8]};
    const double area = ph.dir.x * pr[9] + ph.dir.y * pr[10] + ph.dir.z * pr[11];
    mul_prob(st_sm, ph, clip01(area));
}

// ---------------------------------------------------------------------------
// photon birth: sources and pointing as ops, so a simulation can start from
// nothing but a photon count
// ---------------------------------------------------------------------------
// math/random.py:80-95 RandomArbitraryPdf (sort=True, randomize_in_bin=True); t (global): n, cdf[n],
// sortindex[n], x[n], bin_width[n]
MXB_DEV double arbitrary_pdf(const double* t, double u0, double u1) {
    const int n = (int)__ldg(t);
    const double* cdf = t + 1;
    const double choice = 0. + (__ldg(cdf + n - 1) - 0.) * u0;
    int lo = 0, hi = n;                      // np.searchsorted(cdf, choice): first i with cdf[i] >= choice
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(cdf + mid) < choice) lo = mid + 1; else hi = mid;
    }
    if (lo > n - 1) lo = n - 1;
    const int index = (int)__ldg(t + 1 + n + lo);
    const int below = index > 0 ? index - 1 : n - 1;     // python x[index - 1] wraps around
    return __ldg(t + 1 + 2 * n + below) + __ldg(t + 1 + 3 * n + index) * u1;
}

// math/polarization.py:12-62 polarization_vectors
MXB_DEV V3 polarization_vector(const V3& dir, double angle) {
    const V3 r = normalize(dir);
    const bool conv_x = (fabs(r.x) <= 1e-8) && (fabs(r.z) <= 1e-8);     // np.isclose(., 0.)
    const double rp = conv_x ? r.x : r.y;
    V3 v1{(conv_x ? 1. : 0.) - r.x * rp, (conv_x ? 0. : 1.) - r.y * rp, 0. - r.z * rp};
    v1 = normalize(v1);
    const V3 v2 = cross(r, v1);
    double s, c;
    m_sincos(angle, &s, &c);
    return V3{v1.x * c + v2.x * s, v1.y * c + v2.y * s, v1.z * c + v2.z * s};
}

// source/basesources.py:158-277: p = dt e_mode e_const e_table p_mode p_const p_table sky ra dec
template <typename PP, typename DRAW>
// flags: callable specifications (basesources.py:169,181,203) are evaluated by the host and arrive as columns:
// bit0 energy = energy_in (the core energy plane), bit1 polangle = polangle_in, bit2 the time column is an input
MXB_DEV void op_generate(Photon& ph, PP p, const double* gprog, unsigned long long gid, DRAW draw, int s0, int s1,
                         int s2, int s3, double& time, double& polangle, int flags = 0, double energy_in = 0.0,
                         double polangle_in = 0.0) {
    time = (double)gid * p[0];                       // np.arange(0, T, dt)[i]
    if (flags & 1) {
        ph.energy = energy_in;
    } else if ((int)p[1] == 1) {
        const double u0 = draw(s0), u1 = draw(s1);
        ph.energy = arbitrary_pdf(gprog + (long long)p[3], u0, u1);
    } else {
        ph.energy = 1. * p[2];
    }
    const int pm = (int)p[4];
    if (flags & 2) polangle = polangle_in;
    else if (pm == 1) polangle = 0. + (kTwoPi - 0.) * draw(s2);      // np.random.uniform(0, 2 pi)
    else if (pm == 2) {
        const double u0 = draw(s2), u1 = draw(s3);
        polangle = arbitrary_pdf(gprog + (long long)p[6], u0, u1);
    } else polangle = 1. * p[5];
    ph.prob = 1.;
}

// source/pointing.py:101-177 (+ :180-211 jitter): p = M[9] T[9] north[3] sigma
template <typename PP>
MXB_DEV void op_pointing(Photon& ph, PP p, int flags, double ra_deg, double dec_deg, double polangle, double u_axis,
                         double z_jitter) {
    const double ra = ra_deg * (3.141592653589793 / 180.), dec = dec_deg * (3.141592653589793 / 180.);   // np.deg2rad
    double sr, cr, sd, cd;
    m_sincos(ra, &sr, &cr);
    m_sincos(dec, &sd, &cd);
    const V3 v{cd * cr, cd * sr, sd};
    const V3 o{p[0] * v.x + p[1] * v.y + p[2] * v.z, p[3] * v.x + p[4] * v.y + p[5] * v.z,
               p[6] * v.x + p[7] * v.y + p[8] * v.z};
    const V3 m{-o.x, -o.y, -o.z};
    const V3 d0 = normalize(m);
    PP T = p + 9;
    V3 d{T[0] * d0.x + T[1] * d0.y + T[2] * d0.z, T[3] * d0.x + T[4] * d0.y + T[5] * d0.z,
         T[6] * d0.x + T[7] * d0.y + T[8] * d0.z};
    const V3 north = ld3(p + 18);
    const double proj = d.x * north.x + d.y * north.y + d.z * north.z;
    V3 nin{north.x - d.x * proj, north.y - d.y * proj, north.z - d.z * proj};
    nin = normalize(nin);
    const V3 ein = cross(d, nin);
    double sp, cp;
    m_sincos(polangle, &sp, &cp);
    V3 pol{cp * nin.x + sp * ein.x, cp * nin.y + sp * ein.y, cp * nin.z + sp * ein.z};
    if ((flags & 1) && p[21] > 0.0) {
        double sa, ca;
        sincos_turn(u_axis, &sa, &ca);          // randang = u_axis * 2. * pi
        const V3 ax{0., sa, ca};
        const double ang = 0. + p[21] * z_jitter;
        d = axangle_rotate_T(ax, ang, d);
        pol = axangle_rotate_T(ax, ang, pol);
    }
    ph.dir = d;
    ph.pol = pol;
    ph.unit = false;
}

// source/labSource.py:62-137: p = position[3] R[9] fractional_area
template <typename PP>
MXB_DEV void op_labcone(Photon& ph, PP p, double u_theta, double u_v, double polangle) {
    const double v = 0. + (p[12] - 0.) * u_v;
    const double phi = m_acos(1 - 2 * v);
    double st, ct, sp, cp;
    sincos_turn(u_theta, &st, &ct);            // theta = 0. + (2 pi - 0.) * u_theta
    m_sincos(phi, &sp, &cp);
    const V3 d{ct * sp, st * sp, cp};
    PP R = p + 3;
    ph.dir = V3{R[0] * d.x + R[1] * d.y + R[2] * d.z, R[3] * d.x + R[4] * d.y + R[5] * d.z,
                R[6] * d.x + R[7] * d.y + R[8] * d.z};
    ph.pos = ld3(p);
    ph.pol = polarization_vector(ph.dir, polangle);
    ph.unit = false;
}

// source/labSource.py:13-59: p = pos4d rows 0..2 [12] sourcePos[3]
template <typename PP>
MXB_DEV void op_farlab(Photon& ph, PP p, double u_y, double u_z, double polangle) {
    const double y = -1 + (1 - -1) * u_y, z = -1 + (1 - -1) * u_z;
    ph.pos = V3{p[0] * 0. + p[1] * y + p[2] * z + p[3] * 1., p[4] * 0. + p[5] * y + p[6] * z + p[7] * 1.,
                p[8] * 0. + p[9] * y + p[10] * z + p[11] * 1.};
    ph.dir = V3{ph.pos.x - p[12], ph.pos.y - p[13], ph.pos.z - p[14]};
    ph.pol = polarization_vector(ph.dir, polangle);
    ph.unit = false;
}

// ---------------------------------------------------------------------------
// Parallel containers: facet search with sequential ("last hit wins") semantics
// (simulator.py:42-49 over a Parallel).  H: array header O[3] nbar[3] u[3] v[3]
// u0 v0 inv_cell T2.
// ---------------------------------------------------------------------------
struct ArrayIter {
    int cur, end;
    bool brute;   // candidates are facet indices cur..end-1 themselves (no culling grid, or the steep-ray path)
    bool seg;     // steep ray: every round scans the cells under the ray's footprint (array_scan_segment)
};

// true when `dir` lies inside the cone for which the culling grid is conservative
template <typename HP>
MXB_DEV bool cull_cone_ok(HP H, const V3& dir, bool unit, double& dn) {
    const V3 nb = ld3(H + 3);
    dn = dot(dir, nb);
    const double d2 = (kTrackUnit && unit) ? 1.0 : dot(dir, dir);
    return dn != 0.0 && (d2 - dn * dn) <= H[15] * dn * dn;
}

// start the search of one photon: the candidate range of its culling cell (mode 1) or all F facets
template <typename HP, typename IP>
MXB_DEV void array_open(ArrayIter& it, HP H, IP cell_start, int F, int mode, int nu, int nv, const Photon& ph,
                        bool active, unsigned long long* st_sm) {
    it.brute = true;
    it.seg = false;
    it.cur = 0;
    it.end = active ? F : 0;
    if (mode == 1 && active) {
        double dn;
        const bool ok = cull_cone_ok(H, ph.dir, ph.unit, dn);
        if (!(dn == dn)) {
            it.end = 0;  // NaN direction can never hit (k >= 0 is false)
        } else if (ok) {
            const V3 nb = ld3(H + 3);
            const V3 O = ld3(H);
            const double t = ((O.x - ph.pos.x) * nb.x + (O.y - ph.pos.y) * nb.y + (O.z - ph.pos.z) * nb.z) * fast_rcp(dn);
            const V3 q{ph.pos.x + t * ph.dir.x - O.x, ph.pos.y + t * ph.dir.y - O.y, ph.pos.z + t * ph.dir.z - O.z};
            const double fu = (dot(q, ld3(H + 6)) - H[12]) * H[14];
            const double fv = (dot(q, ld3(H + 9)) - H[13]) * H[14];
            it.brute = false;
            if (fu >= 0.0 && fv >= 0.0 && fu < (double)nu && fv < (double)nv) {
                const int cell = (int)fv * nu + (int)fu;
                it.cur = cell_start.i32(cell);
                it.end = cell_start.i32(cell + 1);
            } else {
                it.cur = it.end = 0;
            }
        } else {
            it.seg = true;   // outside the cone: footprint scan instead of the single cell
            count_status(st_sm, MXB_ST_BRUTE);
        }
    }
}

// Steep rays (outside the cone the single-cell lookup is proven for, or after a second redirection).
// Every facet point lies in the slab |h| <= Hs around the reference plane, so the ray can only hit
// facets while it is inside the slab; the projection of that ray segment is a straight line on the
// grid, and a facet is listed in every cell its footprint touches: scanning the cells of the
// segment's bounding box finds every possible hit, for ANY direction.  Returns the smallest facet
// index > after that the ray hits from (pos, dir), or -1.
#ifdef MXB_OUTLINE_SCAN
#define MXB_SCAN_ATTR __device__ __noinline__
#else
#define MXB_SCAN_ATTR __device__ __forceinline__   // measured (C2, r01): a call site in the search loop costs 18 %
#endif
// MXB_SCAN_UNROLL1 (set by the kernel generator for programs whose array bodies redirect photons more than once, i.e.
// where steep rays are common and every warp walks this scan with one or two lanes): do not unroll the scan loops -
// smaller code in an instruction-cache-bound kernel (measured r02: C3 34.0 -> 32.3 ms; C2, which never scans, loses 1 %
// to the changed register allocation, so it is not the default)
#ifdef MXB_SCAN_UNROLL1
#define MXB_SCAN_PRAGMA _Pragma("unroll 1")
#else
#define MXB_SCAN_PRAGMA
#endif
template <typename BP, typename HP, typename IP>
MXB_SCAN_ATTR int array_scan_segment(BP B, HP H, IP cell_start, IP cand, int rows_off, int stride, int F,
                                               int nu, int nv, V3 pos, V3 dir, int after) {
    const V3 nb = ld3(H + 3), O = ld3(H);
    const double Hs = H[16];
    const V3 rel{pos.x - O.x, pos.y - O.y, pos.z - O.z};
    const double h0 = dot(rel, nb), dn = dot(dir, nb);
    if (!(dn == dn) || !(h0 == h0)) return -1;
    int iu0 = 0, iu1 = nu - 1, iv0 = 0, iv1 = nv - 1;
    if (dn == 0.0) {
        if (fabs(h0) > Hs) return -1;          // parallel to the slab and outside of it
    } else {
        const double ta = div(-Hs - h0, dn), tb = div(Hs - h0, dn);
        const double t_lo = fmax(fmin(ta, tb), 0.0), t_hi = fmax(ta, tb);
        if (t_hi < 0.0) return -1;             // the slab lies behind the photon
        const V3 uu = ld3(H + 6), vv = ld3(H + 9);
        const V3 qa{rel.x + t_lo * dir.x, rel.y + t_lo * dir.y, rel.z + t_lo * dir.z};
        const V3 qb{rel.x + t_hi * dir.x, rel.y + t_hi * dir.y, rel.z + t_hi * dir.z};
        const double fua = (dot(qa, uu) - H[12]) * H[14], fub = (dot(qb, uu) - H[12]) * H[14];
        const double fva = (dot(qa, vv) - H[13]) * H[14], fvb = (dot(qb, vv) - H[13]) * H[14];
        const double ulo = floor(fmin(fua, fub)), uhi = floor(fmax(fua, fub));
        const double vlo = floor(fmin(fva, fvb)), vhi = floor(fmax(fva, fvb));
        if (ulo == ulo && uhi == uhi && vlo == vlo && vhi == vhi) {   // NaN / inf: keep the whole grid
            if (uhi < 0.0 || vhi < 0.0 || ulo > (double)(nu - 1) || vlo > (double)(nv - 1)) return -1;
            iu0 = (int)fmax(ulo, 0.0);
            iv0 = (int)fmax(vlo, 0.0);
            iu1 = (int)fmin(uhi, (double)(nu - 1));
            iv1 = (int)fmin(vhi, (double)(nv - 1));
        }
    }
    V3 ipt;
    double a0, a1;
    if ((long long)(iu1 - iu0 + 1) * (iv1 - iv0 + 1) > 256) {
        // a footprint this long touches most of the array: one ordered pass over the facets is cheaper
        MXB_SCAN_PRAGMA
        for (int j = after + 1; j < F; ++j)
            if (plane_intersect(B + (rows_off + j * stride), pos, dir, false, ipt, a0, a1)) return j;
        return -1;
    }
    int best = 0x7fffffff;
    MXB_SCAN_PRAGMA
    for (int iv = iv0; iv <= iv1; ++iv)
        MXB_SCAN_PRAGMA
        for (int iu = iu0; iu <= iu1; ++iu) {
            const int cell = iv * nu + iu;
            const int k1 = cell_start.i32(cell + 1);
            MXB_SCAN_PRAGMA
            for (int k = cell_start.i32(cell); k < k1; ++k) {
                const int j = cand.i32(k);
                if (j > after && j < best && plane_intersect(B + (rows_off + j * stride), pos, dir, false, ipt, a0, a1)) best = j;
            }
        }
    return best == 0x7fffffff ? -1 : best;
}

// next facet (ascending index) the photon hits from its CURRENT state; row = word offset of its row
template <typename BP, typename HP, typename IP>
MXB_DEV bool array_search(ArrayIter& it, BP B, HP H, IP cell_start, IP cand, int rows_off, int stride, int F, int nu,
                          int nv, Photon& ph, int& row) {
    if (it.seg) {
        if (it.cur >= it.end) return false;
        const int j = array_scan_segment(B, H, cell_start, cand, rows_off, stride, F, nu, nv, ph.pos, ph.dir, it.cur - 1);
        if (j < 0) {
            it.cur = it.end;
            return false;
        }
        row = rows_off + j * stride;
        it.cur = j + 1;
        plane_intersect(B + row, ph.pos, ph.dir, false, ph.ip, ph.l0, ph.l1);
        return true;
    }
    while (it.cur < it.end) {
        const int j = it.brute ? it.cur : cand.i32(it.cur);
        ++it.cur;
        const int r = rows_off + j * stride;
        // results go straight into the photon: after a miss ip / l0 / l1 are dead (only read under hit)
        if (plane_intersect(B + r, ph.pos, ph.dir, false, ph.ip, ph.l0, ph.l1)) {
            row = r;
            return true;
        }
    }
    return false;
}

// after the body, for photons that hit:
// (1) DISJOINTNESS CERTIFICATE (program.py single_hit_successors).  The lowering computes for every pair of facets
//     A < B the smallest tangent (angle to the array's mean normal) a ray needs to get from any point of A to B: the
//     footprints on the mean plane are separated by gap_AB (separating-axis test), the heights differ by at most h_AB,
//     and a ray travels h tan(theta) sideways per height h, so it needs tan(theta) >= gap_AB / h_AB.  Inside the cone
//     tan^2 <= H[17] a photon that leaves facet A can therefore only reach A's short SUCCESSOR list (later facets with
//     a smaller limit: none for the facets of a tiled array, the overlapping diagonal neighbours of a ring-placed
//     one).  lim_mode 1: every list is empty - the photon is DONE with the array; 2: the search continues over the
//     successors only (`limits`: int32 list starts per facet, indexing into the candidate array).  Either way the
//     reference's remaining facets (simulator.py:42-49) would all miss, so they are not tested and a steep diffraction
//     order does not walk the footprint scan.  0: no certificate.
// (2) otherwise, when the body can redirect photons, the culling cone is re-validated for the NEW direction: the cell
//     list covers ONE redirection inside the cone (H t + 2 H t' <= margin); a second hit or a steep new direction
//     falls back to the footprint scan over the remaining facets
template <bool REDIRECTS = true, typename HP, typename LP>
MXB_DEV void array_revalidate(ArrayIter& it, HP H, LP limits, int lim_mode, const Photon& ph, int nhit, int row,
                              int rows_off, int stride, int F, unsigned long long* st_sm) {
    if (!ph.hit) return;
#ifdef MXB_NO_CERT      // A/B switch of the certificate
    lim_mode = 0;
#endif
    if (!lim_mode && (!REDIRECTS || it.brute)) return;
    const V3 nb = ld3(H + 3);
    const double dn = dot(ph.dir, nb);
    const double d2 = (kTrackUnit && ph.unit) ? 1.0 : dot(ph.dir, ph.dir);
    const double c2 = dn * dn, s2 = d2 - c2;          // tan^2 = s2 / c2; NaN fails every comparison below
    if (lim_mode && s2 <= H[17] * c2) {
        if (lim_mode == 2) {      // only this facet's successors can still be hit: an ordinary (short) candidate list
            const int j = (row - rows_off) / stride;
            it.brute = false;
            it.seg = false;
            it.cur = limits.i32(j);
            it.end = limits.i32(j + 1);
        } else {
            it.cur = it.end;      // no other facet can be hit from here
        }
        return;
    }
    if (REDIRECTS && !it.brute) {
        const bool ok = dn != 0.0 && s2 <= H[15] * c2;
        if (nhit >= 2 || (dn == dn && !ok)) {
            const int j = (row - rows_off) / stride;
            it.brute = true;
            it.seg = true;     // footprint scan from the new state, valid for any number of redirections
            it.cur = j + 1;
            it.end = F;
            count_status(st_sm, MXB_ST_BRUTE);
        }
    }
}

}  // namespace mxb
