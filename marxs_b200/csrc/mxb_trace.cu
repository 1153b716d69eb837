// mxb_trace.cu — the fused element-program kernel and the C ABI of libmxb.
//
// One persistent kernel executes a whole lowered Sequence (FlatStacks, Parallel
// facet arrays, detectors) per photon with the photon record held in registers:
// a photon batch crosses HBM once (11 fp64 planes in, 10 out + the diagnostic
// columns the program materialises).  Program, geometry, facet rows and culling
// grids are staged once per CTA into shared memory with a TMA bulk copy
// (cp.async.bulk + mbarrier); CTAs are persistent (grid = SMs x occupancy) and
// grid-stride over photon tiles.  Tensor cores are not used: the work is
// branchy per-photon fp64 (BASELINE.json north_star).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mxb.h"
#include "mxb_device.cuh"

using namespace mxb;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                    \
    do {                                                                                  \
        cudaError_t e_ = (expr);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(MXB_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));   \
    } while (0)

#ifndef MXB_THREADS
#define MXB_THREADS 512
#endif
constexpr int kThreads = MXB_THREADS;   // one persistent CTA per SM: its warps share one staged program
constexpr int kMaxSmemStageBytes = 200 * 1024;

struct TraceParams {
    const double* prog;      // device blob
    int n_ops;
    int stage_words;         // words of the blob staged in smem
    long long n;
    long long id0;
    unsigned long long seed;
    unsigned long long* status;
    MxbColumns cols;
};

// --- TMA bulk copy global -> shared (SASS: UBLKCP), completion on an mbarrier ---
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
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
// program blob access: word offsets into the staged copy (shared memory, 32-bit
// offsets -> LDS) or into the global blob when the program is too large to stage
// ---------------------------------------------------------------------------
extern __shared__ __align__(16) double g_smem[];

template <bool STAGED>
struct PRef {
    const double* g;
    int off;
    __device__ __forceinline__ double operator[](int k) const {
        if (STAGED) return g_smem[off + k];
        return __ldg(g + off + k);
    }
    __device__ __forceinline__ PRef operator+(int k) const { return PRef{g, off + k}; }
    __device__ __forceinline__ double2 ld2(int k) const {  // words 2k, 2k+1 (off must be even)
        if (STAGED) return reinterpret_cast<const double2*>(g_smem + off)[k];
        return __ldg(reinterpret_cast<const double2*>(g + off) + k);
    }
    __device__ __forceinline__ int i32(int k) const {  // packed int32 view of the words at off
        if (STAGED) return reinterpret_cast<const int*>(g_smem)[2 * off + k];
        return __ldg(reinterpret_cast<const int*>(g) + 2 * (long long)off + k);
    }
};

// fused detector image: gp = nx ny sel_lo n_sel ; bin = round half to even like np.round
template <typename PP>
__device__ __forceinline__ void accumulate_image(double* img, PP gp, long long idn, double px, double py,
                                                 double w) {
    const long long nx = (long long)gp[0], ny = (long long)gp[1];
    const long long plane = idn - (long long)gp[2];
    if (plane < 0 || plane >= (long long)gp[3] || !(px == px) || !(py == py) || !(w == w)) return;
    const long long ix = llrint(px), iy = llrint(py);
    if (ix < 0 || iy < 0 || ix >= nx || iy >= ny) return;
    atomicAdd(&img[(plane * ny + iy) * nx + ix], w);
}

// ---------------------------------------------------------------------------
// per-thread photon state
// ---------------------------------------------------------------------------
struct Photon {
    V3 pos, dir, pol;
    double energy, prob;
    V3 ip;          // intersection point of the current element
    double l0, l1;  // local coordinates on the current element
    bool hit;
};

struct Ctx {
    const TraceParams* P;
    long long i;           // photon index in the batch
    bool active;
    bool init_round;            // column-initialising stores are live (see put)
    unsigned long long* st_sm;  // smem status accumulators
};

__device__ __forceinline__ double draw(const Ctx& c, int slot, int kind) {
    const double* inj = c.P->cols.draws[slot];
    if (inj) return inj[c.i];
    return device_draw(c.P->seed, (unsigned long long)(c.P->id0 + c.i), slot, kind);
}

// Output columns are resolved once per CTA: cp[k] = device pointer of the op's k-th column and
// a mode word (bit k: materialised, bit 8+k: this op initialises the column, i.e. stores
// NaN / -1 for photons it does not touch - outside arrays, or in the first search round).
// Per op, `wmask` = the columns this lane has to store (computed once from hit / active /
// init_round); each put is then a predicated global store.
typedef unsigned long long ColEntry;

__device__ __forceinline__ void st_global(double* p, double v) {
    asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_global(long long* p, long long v) {
    asm volatile("st.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ int store_mask(const Ctx& c, int mode, bool hit) {
    const int sel = hit ? 0xff : (c.init_round ? (mode >> 8) : 0);
    return c.active ? (mode & sel) : 0;
}
__device__ __forceinline__ void put(const Ctx& c, int wmask, int k, ColEntry e, bool hit, double v) {
    if (wmask & (1 << k))
        st_global(reinterpret_cast<double*>(e) + c.i, hit ? v : __longlong_as_double(0x7ff8000000000000LL));
}
__device__ __forceinline__ void put_id(const Ctx& c, int wmask, int k, ColEntry e, bool hit, long long v) {
    if (wmask & (1 << k)) st_global(reinterpret_cast<long long*>(e) + c.i, hit ? v : -1LL);
}

// optics/base.py:43-47: probability factors multiply and must lie in [0,1]
__device__ __forceinline__ void mul_prob(const Ctx& c, Photon& ph, double f) {
    if (f < 0.0 || f > 1.0) atomicAdd(&c.st_sm[MXB_ST_PROB_RANGE], 1ULL);
    ph.prob *= f;
}

// ---------------------------------------------------------------------------
// element physics (each only touches photons with ph.hit)
// ---------------------------------------------------------------------------
// mirror.py:53-82  params: P[3] f
template <typename PP>
__device__ __forceinline__ void op_lens(Photon& ph, PP p) {
    const V3 nd = normalize(ph.dir);
    const double f = p[3];
    const V3 t{(p[0] + f * nd.x) - ph.ip.x, (p[1] + f * nd.y) - ph.ip.y, (p[2] + f * nd.z) - ph.ip.z};
    const V3 nd2 = normalize(t);
    ph.pol = parallel_transport(ph.dir, nd2, ph.pol);
    ph.dir = nd2;
}

// scatter.py:49-77  params: center[3] sig_in sig_perp
template <typename PP>
__device__ __forceinline__ void op_rscatter(const Ctx& c, Photon& ph, PP p, int s0, int s1,
                                            double& a, double& b) {
    const V3 radial{ph.pos.x - p[0], ph.pos.y - p[1], ph.pos.z - p[2]};
    const V3 perp = cross(ph.dir, radial);
    V3 out = ph.dir;
    a = 0.0;
    b = 0.0;
    double z0 = 0.0, z1 = 0.0;
    const bool both = (p[3] != 0.0) && (p[4] != 0.0);
    if (both && !c.P->cols.draws[s0] && !c.P->cols.draws[s1]) {
        device_draw_normal_pair(c.P->seed, (unsigned long long)(c.P->id0 + c.i), s0, z0, z1);
    } else {
        if (p[3] != 0.0) z0 = draw(c, s0, 1);
        if (p[4] != 0.0) z1 = draw(c, s1, 1);
    }
    if (p[3] != 0.0) {
        a = p[3] * z0;
        out = axangle_rotate_T(perp, a, ph.dir);
    }
    if (p[4] != 0.0) {
        b = p[4] * z1;
        out = axangle_rotate_T(radial, b, out);
    }
    ph.pol = parallel_transport(ph.dir, out, ph.pol);
    ph.dir = out;
}

// scatter.py:109-145  params: sigma
template <typename PP>
__device__ __forceinline__ void op_gscatter(const Ctx& c, Photon& ph, PP p, int s0, int s1, double& ang) {
    const V3 pdir = normalize(ph.dir);
    const V3 guess = (fabs(pdir.x) < 0.99999) ? V3{1, 0, 0} : V3{0, 1, 0};
    const V3 perp = cross(pdir, guess);
    ang = p[0] * draw(c, s0, 1);
    V3 out = axangle_rotate_T(perp, ang, pdir);
    const double ang2 = draw(c, s1, 0) * 2 * 3.141592653589793;
    out = axangle_rotate_T(pdir, ang2, out);
    ph.pol = parallel_transport(ph.dir, out, ph.pol);
    ph.dir = out;
}

// filter.py:90-94  params: n, x[n], y[n]   (n == 0: constant y[0])
template <typename PP>
__device__ __forceinline__ double filter_value(const Ctx& c, PP p, double energy, int flags) {
    const int n = (int)p[0];
    if (n == 0) return p[1];
    PP xp = p + 1;
    PP fp = p + 1 + n;
    if ((flags & 1) && (energy < xp[0] || energy > xp[n - 1]))
        atomicAdd(&c.st_sm[MXB_ST_FILTER_BOUNDS], 1ULL);
    return interp_clamped(xp, fp, n, energy);
}

// grating.py:12-57, 60-96; mitsnl/catgrating.py:104-144.  Returns order, sets psel.
template <typename PP>
__device__ __forceinline__ double select_order(PP sel, const double* gprog, double u, double energy,
                                               double blaze, double& psel) {
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
        double xq = kHcKevNm / energy;
        xq = fmin(fmax(xq, wk[0]), wk[nw - 1]);
        double yq = fmin(fmax(blaze, tk[0]), tk[nt - 1]);
        const int i = bracket(wk, nw, xq), j = bracket(tk, nt, yq);
        const double tx = (xq - wk[i]) / (wk[i + 1] - wk[i]);
        const double ty = (yq - tk[j]) / (tk[j + 1] - tk[j]);
        const double* t00 = tab + ((long long)i * nt + j) * no;
        const double* t10 = t00 + (long long)nt * no;
        const double* t01 = t00 + no;
        const double* t11 = t10 + no;
        double total = 0.0;
        for (int k = 0; k < no; ++k) {
            const double a00 = __ldg(t00 + k), a10 = __ldg(t10 + k), a01 = __ldg(t01 + k), a11 = __ldg(t11 + k);
            const double f0 = a00 + tx * (a10 - a00);
            const double f1 = a01 + tx * (a11 - a01);
            total = total + (f0 + ty * (f1 - f0));
        }
        psel = total;
        double run = 0.0;
        int oi = 0;
        for (int k = 0; k < no; ++k) {
            const double a00 = __ldg(t00 + k), a10 = __ldg(t10 + k), a01 = __ldg(t01 + k), a11 = __ldg(t11 + k);
            const double f0 = a00 + tx * (a10 - a00);
            const double f1 = a01 + tx * (a11 - a01);
            run = run + (f0 + ty * (f1 - f0));
            if (run / total > u) { oi = k; break; }   // argmax(cumprob > u): first True, 0 if none
        }
        return ord[oi];
    }
}

// grating.py:233-277  params: l[3] dd[3] d blaze0 dblaze ; n = e_x of the geometry
template <typename PP>
__device__ __forceinline__ void op_grating(const Ctx& c, Photon& ph, PP p, PP geom, PP sel,
                                           int flags, int slot, double& order, double& blaze) {
    const V3 pn = normalize(ph.dir);
    const V3 l = ld3(p), dd = ld3(p + 3), n = ld3(geom + 3);
    const double wave = div(kEnergy2Wave, ph.energy);
    const double p_l = dot(pn, l);
    const V3 pp = normalize(V3{pn.x - p_l * l.x, pn.y - p_l * l.y, pn.z - p_l * l.z});
    blaze = acos(clip01(fabs(dot(pp, n))));
    if (flags & 4) blaze = blaze + (p[7] + ph.l0 * p[8]);  // NonParallelCATGrating blaze_angle_modifier
    const double u = draw(c, slot, 0);
    double psel;
    order = select_order(sel, c.P->prog, u, ph.energy, blaze, psel);
    const double p_dd = dot(pn, dd);
    const double sign = (flags & 1) ? ((p_dd < 0.0) ? -1.0 : 1.0) : -1.0;  // CAT: grating.py:298-301
    const double p_d = p_dd + div(sign * order * wave, p[6]);
    const double p_n = sqrt(1. - p_d * p_d - p_l * p_l);
    const double pdn = dot(pn, n);
    double direction = (pdn > 0.0) ? 1.0 : ((pdn < 0.0) ? -1.0 : pdn);  // np.sign
    if (flags & 2) direction = direction * -1;
    const double q = direction * p_n;
    const V3 nd{p_d * dd.x + p_l * l.x + q * n.x, p_d * dd.y + p_l * l.y + q * n.y,
                p_d * dd.z + p_l * l.z + q * n.z};
    ph.pol = parallel_transport(ph.dir, nd, ph.pol);
    ph.dir = nd;
    mul_prob(c, ph, psel);
}

// multiLayerMirror.py:44-91  params: Pinv[9] P[9] ex[3]
template <typename PP>
__device__ __forceinline__ void op_brewster(const Ctx& c, Photon& ph, PP p) {
    const V3 dh = normalize(ph.dir);
    V3 loc{p[0] * dh.x + p[1] * dh.y + p[2] * dh.z, p[3] * dh.x + p[4] * dh.y + p[5] * dh.z,
           p[6] * dh.x + p[7] * dh.y + p[8] * dh.z};
    loc.x = loc.x * -1;
    PP q = p + 9;
    const V3 nd{q[0] * loc.x + q[1] * loc.y + q[2] * loc.z, q[3] * loc.x + q[4] * loc.y + q[5] * loc.z,
                q[6] * loc.x + q[7] * loc.y + q[8] * loc.z};
    const V3 ex = ld3(p + 18);
    V3 v_s = cross(dh, ex);
    const double nvs = sqrt(dot(v_s, v_s));
    v_s = V3{v_s.x / nvs, v_s.y / nvs, v_s.z / nvs};
    const V3 v_p = cross(dh, v_s);
    const double pvs = dot(ph.pol, v_s), pvp = dot(ph.pol, v_p);
    const double Es2 = 1. * (pvs * pvs), Ep2 = 0. * (pvp * pvp);
    const double inten = Es2 + Ep2;
    if (inten > 1.001) atomicAdd(&c.st_sm[MXB_ST_INTENSITY], 1ULL);
    const V3 nvp = cross(nd, v_s);
    const V3 np_{-Es2 * v_s.x + Ep2 * nvp.x, -Es2 * v_s.y + Ep2 * nvp.y, -Es2 * v_s.z + Ep2 * nvp.z};
    const double nn = sqrt(dot(np_, np_));
    ph.pol = V3{np_.x / nn, np_.y / nn, np_.z / nn};
    ph.dir = nd;
    mul_prob(c, ph, clip01(inten));
}

// multiLayerMirror.py:132-170  params: Ly n_refl n_pol xs[nr] peak_lambda[nr] peak[nr] fwhm[nr] pol_e[np] pol[np]
template <typename PP>
__device__ __forceinline__ void op_mleff(const Ctx& c, Photon& ph, PP p) {
    const double Ly = p[0];
    const int nr = (int)p[1], npol = (int)p[2];
    PP xs = p + 3;
    PP pl = xs + nr;
    PP pk = pl + nr;
    PP fw = pk + nr;
    PP pe = fw + nr;
    PP pf = pe + npol;
    const double wavelength = kHcMultilayer / ph.energy;
    const double tested = interp_clamped(pe, pf, npol, ph.energy);
    const double local_x = ph.l0 / Ly;
    const double peak_w = interp_clamped(xs, pl, nr, local_x);
    const double max_refl = interp_clamped(xs, pk, nr, local_x) / tested;
    const double spread = interp_clamped(xs, fw, nr, local_x);
    const double c2 = (spread * spread) / (8. * 0.6931471805599453);
    double refl = 0.0;
    if (c2 != 0.0) {
        const double dw = wavelength - peak_w;
        refl = max_refl * exp(-(dw * dw) / (2 * c2));
    }
    mul_prob(c, ph, refl / 100);
}

// ---------------------------------------------------------------------------
// the fused program interpreter
// ---------------------------------------------------------------------------
struct OpHot {   // one LDS.128 per dispatch
    int type;    // op code | kRelBit when the element params live in the current facet row
    int flags;
    int poff;    // word offset of the element params (relative to the facet row when kRelBit)
    int pg;      // word offset of the global block (geometry, selector table, array header ...)
};
struct OpCold {
    ColEntry cp[8];   // resolved output columns (not for ARRAY_BEGIN / LOADHIT, whose c[] are plain ints)
    int c[8];
    int s0, s1, w14, w15;
    int mode;         // bit k: column k materialised, bit 8+k: initialised by this op
    int pad[3];
};
constexpr int kRelBit = 1 << 20;
constexpr int kCommitBit = 256;     // flags: a COMMIT is folded into this op (columns c5..c7)
constexpr int kCommitRowId = 512;   // flags: its id_num comes from the facet row (word w15)

template <bool STAGED>
__global__ void __launch_bounds__(kThreads, 1)
mxb_trace_kernel(const __grid_constant__ TraceParams P) {
    __shared__ OpHot oph[MXB_MAX_OPS + 1];
    __shared__ OpCold opc[MXB_MAX_OPS];
    __shared__ unsigned long long st_sm[MXB_STATUS_WORDS];
    __shared__ unsigned int hits_w[kThreads / 32][MXB_MAX_OPS];   // per-warp hit counters (no atomics)
    __shared__ __align__(8) uint64_t bar;
    typedef PRef<STAGED> Ref;

    const int tid = threadIdx.x;
    // ---- stage the program: TMA bulk copy of the hot part of the blob ----
    if (STAGED) {
        if (tid == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)P.stage_words * 8u;
            mbar_expect_tx(&bar, bytes);
            uint32_t off = 0;
            while (off < bytes) {
                const uint32_t chunk = min(bytes - off, 32768u);
                bulk_g2s(reinterpret_cast<char*>(g_smem) + off, reinterpret_cast<const char*>(P.prog) + off,
                         chunk, &bar);
                off += chunk;
            }
        }
    }
    for (int k = tid; k < MXB_STATUS_WORDS; k += kThreads) st_sm[k] = 0ULL;
    for (int k = tid; k < (kThreads / 32) * MXB_MAX_OPS; k += kThreads) (&hits_w[0][0])[k] = 0u;
    if (STAGED) mbar_wait(&bar, 0);
    const Ref B{P.prog, 0};
    for (int k = tid; k < P.n_ops; k += kThreads) {
        const Ref w = B + (MXB_HEADER_WORDS + k * MXB_OP_WORDS);
        const int pg = (int)w[2], pf = (int)w[3];
        OpHot h;
        h.type = (int)w[0] | (pf >= 0 ? kRelBit : 0);
        h.flags = (int)w[1];
        h.poff = pf >= 0 ? pf : (pg >= 0 ? pg : 0);
        h.pg = pg >= 0 ? pg : 0;
        oph[k] = h;
        OpCold c;
        const int otype = (int)w[0];
        c.mode = 0;
        c.pad[0] = c.pad[1] = c.pad[2] = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = (int)w[4 + j];
            c.c[j] = col;
            ColEntry e = 0ULL;
            if (col >= 0 && otype != MXB_OP_ARRAY_BEGIN) {
                const int idx = col >= MXB_COL_INIT ? col - MXB_COL_INIT : col;
                const bool is_id = (otype == MXB_OP_COMMIT && j == 2) || (otype != MXB_OP_COMMIT && j == 7 && (((int)w[1]) & kCommitBit));
                if (is_id ? idx < MXB_MAX_I64_COLS : idx < MXB_MAX_F64_COLS) {
                    const unsigned long long p = is_id ? (unsigned long long)P.cols.i64[idx]
                                                       : (unsigned long long)P.cols.f64[idx];
                    e = p;
                    if (p) c.mode |= (1 << j) | (col >= MXB_COL_INIT ? (256 << j) : 0);
                }
            }
            c.cp[j] = e;
        }
        c.s0 = (int)w[12];
        c.s1 = (int)w[13];
        c.w14 = (int)w[14];
        c.w15 = (int)w[15];
        opc[k] = c;
    }
    __syncthreads();

    Ctx ctx;
    ctx.P = &P;
    ctx.st_sm = st_sm;
    const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
    const long long stride = (long long)gridDim.x * kThreads;
    const long long n_round = ((P.n + kThreads - 1) / kThreads) * kThreads;  // keep warps whole
    const bool lane0 = (tid & 31) == 0;
    unsigned int* my_hits = hits_w[tid >> 5];

    for (long long i = (long long)blockIdx.x * kThreads + tid; i < n_round; i += stride) {
        ctx.i = i;
        ctx.active = i < P.n;
        Photon ph;
        if (ctx.active) {
            const MxbColumns& C = P.cols;
            ph.pos = V3{C.f64[0][i], C.f64[1][i], C.f64[2][i]};
            ph.dir = V3{C.f64[3][i], C.f64[4][i], C.f64[5][i]};
            ph.pol = V3{C.f64[6][i], C.f64[7][i], C.f64[8][i]};
            ph.energy = C.f64[9][i];
            ph.prob = C.f64[10][i];
        } else {
            ph.pos = ph.dir = ph.pol = V3{kNaN, kNaN, kNaN};
            ph.energy = ph.prob = kNaN;
        }
        ph.hit = false;
        ph.ip = V3{kNaN, kNaN, kNaN};
        ph.l0 = ph.l1 = kNaN;

        // array iteration state
        int arr_cur = 0, arr_end = 0, arr_nhit = 0, arr_pc = -1;
        bool arr_brute = false;
        ctx.init_round = true;
        int row = 0;    // word offset of the current facet row (0 = the blob itself outside arrays)
        int geom = 0;   // word offset of the current geometry block

        // leave the array whose ARRAY_BEGIN is op `bpc`
        auto array_exit = [&]() {
            if (arr_nhit >= 2) atomicAdd(&st_sm[MXB_ST_MULTI_HIT], 1ULL);
            arr_pc = -1;
            row = 0;
            ph.hit = false;
            ctx.init_round = true;
        };

        int pc = 0;
        OpHot op_next = oph[0];
        while (pc < P.n_ops) {
            const OpHot op = op_next;
            const int pc_in = pc;
            op_next = oph[pc + 1];      // prefetch (oph has MXB_MAX_OPS + 1 entries); re-read below if control jumps
            const Ref pr = B + (((op.type & kRelBit) ? row : 0) + op.poff);
            switch (op.type & (kRelBit - 1)) {
            case MXB_OP_PLANE: {
                geom = op.pg;
                ph.hit = plane_intersect(B + geom, ph.pos, ph.dir, op.flags & 1, ph.ip, ph.l0, ph.l1) && ctx.active;
                const unsigned m = __ballot_sync(0xffffffffu, ph.hit);
                if (lane0) my_hits[pc] += __popc(m);
                break;
            }
            case MXB_OP_LOADHIT: {
                const OpCold& c = opc[pc];
                geom = op.pg;
                ph.hit = false;
                if (ctx.active) {
                    const MxbColumns& C = P.cols;
                    ph.hit = C.f64[c.c[0]][i] != 0.0;
                    ph.ip = V3{C.f64[c.c[1]][i], C.f64[c.c[2]][i], C.f64[c.c[3]][i]};
                    ph.l0 = C.f64[c.c[4]][i];
                    ph.l1 = C.f64[c.c[5]][i];   // (input columns: plain indices)
                }
                const unsigned m = __ballot_sync(0xffffffffu, ph.hit);
                if (lane0) my_hits[pc] += __popc(m);
                break;
            }
            case MXB_OP_COMMIT: {
                const OpCold& c = opc[pc];
                const int wm = store_mask(ctx, c.mode, ph.hit);
                put(ctx, wm, 0, c.cp[0], ph.hit, ph.l0);
                put(ctx, wm, 1, c.cp[1], ph.hit, ph.l1);
                if (c.mode & 4) {
                    const long long idn = (op.flags & 1) ? (long long)(B + row)[c.w15] : (long long)c.w14;
                    put_id(ctx, wm, 2, c.cp[2], ph.hit, idn);
                }
                if (ph.hit) ph.pos = ph.ip;
                break;
            }
            case MXB_OP_BAFFLE: {
                if (ph.hit) ph.pos = ph.ip; else ph.prob = 0.0;
                break;
            }
            case MXB_OP_LENS: {
                if (ph.hit) op_lens(ph, pr);
                break;
            }
            case MXB_OP_RSCATTER: {
                const OpCold& c = opc[pc];
                const int wm = store_mask(ctx, c.mode, ph.hit);
                double a = 0, b = 0;
                if (ph.hit) op_rscatter(ctx, ph, pr, c.s0, c.s1, a, b);
                put(ctx, wm, 0, c.cp[0], ph.hit, a);
                put(ctx, wm, 1, c.cp[1], ph.hit, b);
                break;
            }
            case MXB_OP_GSCATTER: {
                const OpCold& c = opc[pc];
                const int wm = store_mask(ctx, c.mode, ph.hit);
                double a = 0;
                if (ph.hit) op_gscatter(ctx, ph, pr, c.s0, c.s1, a);
                put(ctx, wm, 0, c.cp[0], ph.hit, a);
                break;
            }
            case MXB_OP_FILTER: {
                if (ph.hit) mul_prob(ctx, ph, filter_value(ctx, pr, ph.energy, op.flags));
                break;
            }
            case MXB_OP_GFILTER: {
                if (ctx.active) mul_prob(ctx, ph, filter_value(ctx, pr, ph.energy, op.flags));
                break;
            }
            case MXB_OP_GRATING: {
                const OpCold& c = opc[pc];
                const int wm = store_mask(ctx, c.mode, ph.hit);
                double order = 0, blaze = 0;
                if (ph.hit) op_grating(ctx, ph, pr, B + geom, B + op.pg, op.flags, c.s0, order, blaze);
                put(ctx, wm, 0, c.cp[0], ph.hit, order);
                put(ctx, wm, 1, c.cp[1], ph.hit, blaze);
                break;
            }
            case MXB_OP_DETPIX: {
                // detector.py:73-75; pr: pixsize cp0 cp1; optional fused image: pg: nx ny sel_lo n_sel, s0 image slot
                const OpCold& c = opc[pc];
                const int wm = store_mask(ctx, c.mode, ph.hit);
                const double px = div(ph.l0, pr[0]) + pr[1];
                const double py = div(ph.l1, pr[0]) + pr[2];
                put(ctx, wm, 0, c.cp[0], ph.hit, px);
                put(ctx, wm, 1, c.cp[1], ph.hit, py);
                if (c.s0 >= 0 && ph.hit) {
                    const Ref gp = B + op.pg;
                    const long long idn = (op.flags & 1) ? (long long)(B + row)[c.w15] : (long long)c.w14;
                    accumulate_image(P.cols.f64[c.s0], gp, idn, px, py, ph.prob);
                }
                break;
            }
            case MXB_OP_ACIS: {
                // det_acis.py:31-58 ; per-facet pr: pixsize cp0 cp1 sh ct st ox oy ; global: f pixrad odet0 odet1 cosr sinr
                const OpCold& c = opc[pc];
                const int wm = store_mask(ctx, c.mode, ph.hit);
                const Ref gp = B + op.pg;
                const double chipx = div(ph.l0, pr[0]) + pr[1] + 1;
                const double chipy = div(ph.l1, pr[0]) + pr[2] + 1;
                const double tx = pr[3] * (pr[4] * (chipx - 0.5) + pr[5] * (chipy - 0.5)) + pr[6];
                const double ty = pr[3] * (-pr[5] * (chipx - 0.5) + pr[4] * (chipy - 0.5)) + pr[7];
                const double mn0 = ph.ip.x - gp[0];
                const double x = div(div(ph.ip.y, mn0), gp[1]);
                const double y = div(div(ph.ip.z, mn0), gp[1]);
                put(ctx, wm, 0, c.cp[0], ph.hit, chipx);
                put(ctx, wm, 1, c.cp[1], ph.hit, chipy);
                put(ctx, wm, 2, c.cp[2], ph.hit, tx);
                put(ctx, wm, 3, c.cp[3], ph.hit, ty);
                put(ctx, wm, 4, c.cp[4], ph.hit, gp[2] - x);
                put(ctx, wm, 5, c.cp[5], ph.hit, gp[3] + y);
                put(ctx, wm, 6, c.cp[6], ph.hit, gp[2] - x * gp[4] + y * gp[5]);
                put(ctx, wm, 7, c.cp[7], ph.hit, gp[3] + x * gp[5] + y * gp[4]);
                if (c.s0 >= 0 && ph.hit) {
                    // fused detector image (chip pixel convention is 1-based: det_acis.py:33-34)
                    const long long idn = (long long)(B + row)[c.w15];
                    accumulate_image(P.cols.f64[c.s0], gp + 6, idn, chipx - 1.0, chipy - 1.0, ph.prob);
                }
                break;
            }
            case MXB_OP_BREWSTER: {
                if (ph.hit) op_brewster(ctx, ph, pr);
                break;
            }
            case MXB_OP_MLEFF: {
                if (ph.hit) op_mleff(ctx, ph, pr);
                break;
            }
            case MXB_OP_APERTURE: {
                // aperture.py:42-78: params c[3] vy[3] vz[3] nex[3] phi0 dphi rin2 cum_lo cum_hi
                const OpCold& c = opc[pc];
                bool sel = ctx.active;
                if (c.w14 >= 0 && sel) {  // MultiAperture :201-218: injected aperture id, or area-weighted draw
                    const double a = draw(ctx, c.w14, 0);
                    if (P.cols.draws[c.w14]) sel = ((long long)a == (long long)c.w15);
                    else sel = (a >= pr[15] && a < pr[16]);
                }
                ph.hit = sel;
                if (sel) {
                    double x, y;
                    const double u0 = draw(ctx, c.s0, 0), u1 = draw(ctx, c.s1, 0);
                    if (op.flags & 1) {  // CircleAperture :138-146
                        const double phi = pr[12] + pr[13] * u0;
                        const double r = sqrt(pr[14] + (1. - pr[14]) * u1);
                        double sn, cs;
                        sincos(phi, &sn, &cs);
                        x = r * cs;
                        y = r * sn;
                    } else {  // RectangleAperture :92-95
                        x = u0 * 2. - 1.;
                        y = u1 * 2. - 1.;
                    }
                    ph.l0 = x;
                    ph.l1 = y;
                    ph.ip = V3{pr[0] + x * pr[3] + y * pr[6], pr[1] + x * pr[4] + y * pr[7],
                               pr[2] + x * pr[5] + y * pr[8]};
                    const double area = ph.dir.x * pr[9] + ph.dir.y * pr[10] + ph.dir.z * pr[11];
                    mul_prob(ctx, ph, clip01(area));
                }
                const unsigned m = __ballot_sync(0xffffffffu, ph.hit);
                if (lane0) my_hits[pc] += __popc(m);
                break;
            }
            case MXB_OP_PROPAGATE: {
                ph.pos = V3{ph.pos.x + pr[0] * ph.dir.x, ph.pos.y + pr[0] * ph.dir.y, ph.pos.z + pr[0] * ph.dir.z};
                break;
            }
            case MXB_OP_ARRAY_BEGIN: {
                // op ints: c0 F, c1 row stride, c2 rows offset, c3 mode (1 = culling grid), c4 nu, c5 nv,
                //          c6 cell_start offset (int32), c7 candidate offset (int32), s0 n_init, s1 init offset (int32)
                // pg doubles: O[3] nbar[3] u[3] v[3] u0 v0 inv_cell T2
                const OpCold& c = opc[pc];
                const Ref H = B + op.pg;
                if (arr_pc != pc) {  // first round of this array for this photon
                    arr_pc = pc;
                    ctx.init_round = true;    // the body's first pass initialises the columns it creates
                    arr_nhit = 0;
                    arr_brute = true;
                    arr_cur = 0;
                    arr_end = ctx.active ? c.c[0] : 0;
                    if (c.c[3] == 1 && ctx.active) {
                        const V3 nb = ld3(H + 3);
                        const double dn = dot(ph.dir, nb);
                        const double d2 = dot(ph.dir, ph.dir);
                        if (!(dn == dn)) {
                            arr_end = 0;  // NaN direction can never hit (k >= 0 is false)
                        } else if (dn != 0.0 && (d2 - dn * dn) <= H[15] * dn * dn) {
                            const V3 O = ld3(H);
                            const double t = ((O.x - ph.pos.x) * nb.x + (O.y - ph.pos.y) * nb.y + (O.z - ph.pos.z) * nb.z) * fast_rcp(dn);
                            const V3 q{ph.pos.x + t * ph.dir.x - O.x, ph.pos.y + t * ph.dir.y - O.y, ph.pos.z + t * ph.dir.z - O.z};
                            const double fu = (dot(q, ld3(H + 6)) - H[12]) * H[14];
                            const double fv = (dot(q, ld3(H + 9)) - H[13]) * H[14];
                            arr_brute = false;
                            if (fu >= 0.0 && fv >= 0.0 && fu < (double)c.c[4] && fv < (double)c.c[5]) {
                                const int cell = (int)fv * c.c[4] + (int)fu;
                                const Ref cs = B + c.c[6];
                                arr_cur = cs.i32(cell);
                                arr_end = cs.i32(cell + 1);
                            } else {
                                arr_cur = arr_end = 0;
                            }
                        } else {
                            atomicAdd(&st_sm[MXB_ST_BRUTE], 1ULL);
                        }
                    }
                }
                // search the next facet (ascending index) this photon hits from its CURRENT state
                bool found = false;
                const Ref cand = B + c.c[7];
                while (arr_cur < arr_end) {
                    const int j = arr_brute ? arr_cur : cand.i32(arr_cur);
                    ++arr_cur;
                    const int r = c.c[2] + j * c.c[1];
                    V3 ipt;
                    double a0, a1;
                    if (plane_intersect(B + r, ph.pos, ph.dir, false, ipt, a0, a1)) {
                        found = true;
                        row = r;
                        geom = r;
                        ph.ip = ipt;
                        ph.l0 = a0;
                        ph.l1 = a1;
                        break;
                    }
                }
                ph.hit = found;
                arr_nhit += found ? 1 : 0;
                const unsigned m = __ballot_sync(0xffffffffu, found);
                if (m == 0u) {
                    // no lane found a facet: leave the array, skipping the body (c.w14 = pc of ARRAY_END)
                    if (ctx.init_round && ctx.active) {
                        // the body never ran for this warp: initialise the columns it would have created
                        const Ref init = B + c.s1;
                        for (int k = 0; k < c.s0; ++k) {
                            const int cr = init.i32(k);
                            if (cr >= 0) P.cols.f64[cr][i] = kNaN; else P.cols.i64[-cr - 2][i] = -1LL;
                        }
                    }
                    array_exit();
                    pc = c.w14;
                } else if (lane0) {
                    my_hits[pc] += __popc(m);
                }
                break;
            }
            case MXB_OP_ARRAY_END: {
                // after the body: photons that hit re-validate the culling cone for their NEW direction
                const int bpc = arr_pc;
                const OpCold& c = opc[bpc];
                if (ph.hit && !arr_brute) {
                    const Ref H = B + oph[bpc].pg;
                    const V3 nb = ld3(H + 3);
                    const double dn = dot(ph.dir, nb);
                    const double d2 = dot(ph.dir, ph.dir);
                    // the cell list covers ONE redirection inside the cone (H t + 2 H t' <= margin);
                    // a second hit or a steep new direction falls back to brute force
                    if (arr_nhit >= 2 || (dn == dn && !(dn != 0.0 && (d2 - dn * dn) <= H[15] * dn * dn))) {
                        const int j = (row - c.c[2]) / c.c[1];
                        arr_brute = true;
                        arr_cur = j + 1;
                        arr_end = c.c[0];
                        atomicAdd(&st_sm[MXB_ST_BRUTE], 1ULL);
                    }
                }
                // another search round only if some lane that hit still has facets to test
                ctx.init_round = false;               // later rounds only store for photons that hit
                if (__any_sync(0xffffffffu, ph.hit && arr_cur < arr_end)) {
                    if (!ph.hit) arr_cur = arr_end;   // lanes that found nothing are done with this array
                    pc = bpc - 1;                     // -> ARRAY_BEGIN (pc++ below)
                } else {
                    array_exit();
                }
                break;
            }
            default:
                break;
            }
            if (op.flags & kCommitBit) {
                // optics/base.py:201-209 folded into the element's op: c5,c6 loc-coos columns, c7 id column
                const OpCold& c = opc[pc];
                const int wm = store_mask(ctx, c.mode, ph.hit);
                put(ctx, wm, 5, c.cp[5], ph.hit, ph.l0);
                put(ctx, wm, 6, c.cp[6], ph.hit, ph.l1);
                if (c.mode & 128) {
                    const long long idn = (op.flags & kCommitRowId) ? (long long)(B + row)[c.w15] : (long long)c.w14;
                    put_id(ctx, wm, 7, c.cp[7], ph.hit, idn);
                }
                if (ph.hit) ph.pos = ph.ip;
            }
            ++pc;
            if (pc != pc_in + 1) op_next = oph[pc];
        }

        if (ctx.active) {
            const MxbColumns& C = P.cols;
            C.f64[0][i] = ph.pos.x; C.f64[1][i] = ph.pos.y; C.f64[2][i] = ph.pos.z;
            C.f64[3][i] = ph.dir.x; C.f64[4][i] = ph.dir.y; C.f64[5][i] = ph.dir.z;
            C.f64[6][i] = ph.pol.x; C.f64[7][i] = ph.pol.y; C.f64[8][i] = ph.pol.z;
            C.f64[10][i] = ph.prob;
        }
    }

    __syncthreads();
    for (int k = tid; k < MXB_MAX_OPS; k += kThreads) {
        unsigned long long t = 0ULL;
        for (int w = 0; w < kThreads / 32; ++w) t += hits_w[w][k];
        if (t) atomicAdd(&P.status[MXB_ST_OPHITS + k], t);
    }
    for (int k = tid; k < MXB_ST_OPHITS; k += kThreads)
        if (st_sm[k]) atomicAdd(&P.status[k], st_sm[k]);
}

}  // namespace

// ===========================================================================
// standalone kernels: Geometry.intersect, parallel_transport, detector image
// ===========================================================================
namespace {

struct Geom14 {
    double g[14];
    __device__ __forceinline__ double2 ld2(int k) const { return make_double2(g[2 * k], g[2 * k + 1]); }
};

__global__ void __launch_bounds__(256)
plane_intersect_kernel(const __grid_constant__ Geom14 G, int circular, const double* dx, const double* dy,
                       const double* dz, const double* px, const double* py, const double* pz,
                       unsigned char* hit, double* ix, double* iy, double* iz, double* l0, double* l1,
                       long long n) {
    const double kNaN = __longlong_as_double(0x7ff8000000000000LL);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        V3 ip;
        double a0, a1;
        bool rect;
        const bool h = plane_intersect(G, V3{px[i], py[i], pz[i]}, V3{dx[i], dy[i], dz[i]}, circular != 0,
                                       ip, a0, a1, &rect);
        hit[i] = h ? 1 : 0;
        // geometry.py:254-259 NaN fill uses the rectangle mask (CircularHole narrows it afterwards)
        ix[i] = rect ? ip.x : kNaN;
        iy[i] = rect ? ip.y : kNaN;
        iz[i] = rect ? ip.z : kNaN;
        l0[i] = rect ? a0 : kNaN;
        l1[i] = rect ? a1 : kNaN;
    }
}

__global__ void __launch_bounds__(256)
parallel_transport_kernel(const double* ax, const double* ay, const double* az, const double* bx,
                          const double* by, const double* bz, const double* px, const double* py,
                          const double* pz, double* ox, double* oy, double* oz, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const V3 r = parallel_transport(V3{ax[i], ay[i], az[i]}, V3{bx[i], by[i], bz[i]},
                                        V3{px[i], py[i], pz[i]});
        ox[i] = r.x;
        oy[i] = r.y;
        oz[i] = r.z;
    }
}

__global__ void __launch_bounds__(256)
hist2d_kernel(const double* x, const double* y, const double* w, const long long* sel, long long sel_lo,
              int n_sel, double x0, double y0, long long n, int nx, int ny, double* img,
              unsigned long long* counts) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        long long plane = 0;
        if (sel) {
            plane = sel[i] - sel_lo;
            if (plane < 0 || plane >= n_sel) continue;
        }
        const double xv = x[i] - x0, yv = y[i] - y0;
        if (!(xv == xv) || !(yv == yv)) continue;
        const long long ix = llrint(xv), iy = llrint(yv);   // np.round: half to even
        if (ix < 0 || iy < 0 || ix >= nx || iy >= ny) continue;
        const long long bin = (plane * ny + iy) * nx + ix;
        const double wv = w ? w[i] : 1.0;
        if (img && wv == wv) atomicAdd(&img[bin], wv);
        if (counts) atomicAdd(&counts[bin], 1ULL);
    }
}

int sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
    }
    return sms;
}

int grid_for(long long n, int threads, int per_sm) {
    long long blocks = (n + threads - 1) / threads;
    long long cap = (long long)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int validate_program(const double* h, size_t words, int* n_ops, int* stage_words) {
    if (!h || words < MXB_HEADER_WORDS) return fail(MXB_EINVAL, "program: missing header");
    if ((long long)h[0] != MXB_MAGIC) return fail(MXB_EINVAL, "program: bad magic");
    if ((int)h[1] != MXB_ABI_VERSION) return fail(MXB_EINVAL, "program: ABI version mismatch");
    *n_ops = (int)h[2];
    *stage_words = (int)h[4];
    if (*n_ops < 0 || *n_ops > MXB_MAX_OPS) return fail(MXB_EINVAL, "program: too many ops");
    if ((size_t)h[3] != words) return fail(MXB_EINVAL, "program: size field does not match");
    if (*stage_words < MXB_HEADER_WORDS + *n_ops * MXB_OP_WORDS || (size_t)*stage_words > words)
        return fail(MXB_EINVAL, "program: bad staging size");
    if (*stage_words % 2) return fail(MXB_EINVAL, "program: staging size must be a multiple of 16 bytes");
    return MXB_OK;
}

int launch_trace(const double* prog_dev, int n_ops, int stage_words, const MxbColumns* cols, int64_t n,
                 int64_t id0, uint64_t seed, unsigned long long* status_dev, cudaStream_t stream) {
    TraceParams P;
    P.prog = prog_dev;
    P.n_ops = n_ops;
    P.stage_words = stage_words;
    P.n = n;
    P.id0 = id0;
    P.seed = seed;
    P.status = status_dev;
    P.cols = *cols;
    const size_t stage_bytes = (size_t)stage_words * 8;
    const int grid = grid_for(n, kThreads, 1);
    if (stage_bytes <= (size_t)kMaxSmemStageBytes) {
        static bool attr_set = false;
        if (!attr_set) {
            CUDA_TRY(cudaFuncSetAttribute(mxb_trace_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          kMaxSmemStageBytes));
            attr_set = true;
        }
        mxb_trace_kernel<true><<<grid, kThreads, stage_bytes, stream>>>(P);
    } else {
        mxb_trace_kernel<false><<<grid, kThreads, 0, stream>>>(P);
    }
    CUDA_TRY(cudaGetLastError());
    return MXB_OK;
}

// cached staging state of mxb_trace_host (one per host thread)
struct HostStage {
    int dev = -1;
    cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[3] = {}, ev_k[3] = {}, ev_out[3] = {};
    double* dprog = nullptr;
    size_t prog_bytes = 0;
    unsigned long long* dstatus = nullptr;
    char* dbuf[3] = {nullptr, nullptr, nullptr};
    size_t buf_bytes = 0;
    void release() {
        if (s_in) {
            for (int b = 0; b < 3; ++b) {
                cudaEventDestroy(ev_in[b]);
                cudaEventDestroy(ev_k[b]);
                cudaEventDestroy(ev_out[b]);
            }
            cudaStreamDestroy(s_in);
            cudaStreamDestroy(s_k);
            cudaStreamDestroy(s_out);
        }
        for (int b = 0; b < 3; ++b) if (dbuf[b]) cudaFree(dbuf[b]);
        if (dprog) cudaFree(dprog);
        if (dstatus) cudaFree(dstatus);
        *this = HostStage();
    }
};
HostStage& host_stage() {
    static thread_local HostStage s;
    return s;
}

}  // namespace

extern "C" {

int mxb_version(void) { return MXB_ABI_VERSION; }

const char* mxb_build_info(void) {
#ifdef MXB_FAST
    return "libmxb sm_100a fp64 fast (fmad=on, reciprocal normalisation)";
#else
    return "libmxb sm_100a fp64 strict (fmad=off, IEEE division: bit-parity build)";
#endif
}

const char* mxb_last_error(void) { return g_last_error.c_str(); }

void mxb_host_release(void) { host_stage().release(); }

int mxb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mxb_trace(const double* prog_dev, size_t prog_words, const double* prog_host, const MxbColumns* cols,
              int64_t n, int64_t photon_id0, uint64_t seed, unsigned long long* status_dev, void* stream) {
    int n_ops = 0, stage_words = 0;
    const int rc = validate_program(prog_host, prog_words, &n_ops, &stage_words);
    if (rc) return rc;
    if (!prog_dev || !cols || !status_dev) return fail(MXB_EINVAL, "mxb_trace: null pointer");
    if (n < 0) return fail(MXB_EINVAL, "mxb_trace: negative photon count");
    if (n == 0) return MXB_OK;
    for (int k = 0; k <= MXB_COL_PROB; ++k)
        if (!cols->f64[k]) return fail(MXB_EINVAL, "mxb_trace: core photon column missing");
    return launch_trace(prog_dev, n_ops, stage_words, cols, n, photon_id0, seed, status_dev,
                        (cudaStream_t)stream);
}

int mxb_plane_intersect(const double* geom14_host, int circular, const double* const dir[3],
                        const double* const pos[3], unsigned char* hit, double* const interpos[3],
                        double* const loc[2], int64_t n, void* stream) {
    if (!geom14_host || !dir || !pos || !hit || !interpos || !loc) return fail(MXB_EINVAL, "null pointer");
    if (n <= 0) return n == 0 ? MXB_OK : fail(MXB_EINVAL, "negative n");
    Geom14 G;
    memcpy(G.g, geom14_host, sizeof(G.g));
    plane_intersect_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        G, circular, dir[0], dir[1], dir[2], pos[0], pos[1], pos[2], hit, interpos[0], interpos[1],
        interpos[2], loc[0], loc[1], n);
    CUDA_TRY(cudaGetLastError());
    return MXB_OK;
}

int mxb_parallel_transport(const double* const dir_old[3], const double* const dir_new[3],
                           const double* const pol_old[3], double* const pol_new[3], int64_t n,
                           void* stream) {
    if (!dir_old || !dir_new || !pol_old || !pol_new) return fail(MXB_EINVAL, "null pointer");
    if (n <= 0) return n == 0 ? MXB_OK : fail(MXB_EINVAL, "negative n");
    parallel_transport_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        dir_old[0], dir_old[1], dir_old[2], dir_new[0], dir_new[1], dir_new[2], pol_old[0], pol_old[1],
        pol_old[2], pol_new[0], pol_new[1], pol_new[2], n);
    CUDA_TRY(cudaGetLastError());
    return MXB_OK;
}

int mxb_hist2d(const double* x, const double* y, const double* w, const long long* sel, long long sel_lo,
               int n_sel, double x0, double y0, int64_t n, int nx, int ny, double* img,
               unsigned long long* counts, void* stream) {
    if (!x || !y || (!img && !counts) || nx <= 0 || ny <= 0 || n_sel <= 0)
        return fail(MXB_EINVAL, "mxb_hist2d: bad argument");
    if (n <= 0) return n == 0 ? MXB_OK : fail(MXB_EINVAL, "negative n");
    hist2d_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, y, w, sel, sel_lo, n_sel, x0, y0, n,
                                                                         nx, ny, img, counts);
    CUDA_TRY(cudaGetLastError());
    return MXB_OK;
}

// ---------------------------------------------------------------------------
// host-buffer entry: chunked H2D -> kernel -> D2H pipeline on three streams
// ---------------------------------------------------------------------------
int mxb_trace_host(const double* prog_host, size_t prog_words, const MxbColumns* host_in,
                   const MxbColumns* host_out, int64_t n, int64_t chunk, int64_t photon_id0,
                   uint64_t seed, unsigned long long* status_host) {
    int n_ops = 0, stage_words = 0;
    int rc = validate_program(prog_host, prog_words, &n_ops, &stage_words);
    if (rc) return rc;
    if (!host_in || !host_out || !status_host) return fail(MXB_EINVAL, "mxb_trace_host: null pointer");
    if (n < 0) return fail(MXB_EINVAL, "negative n");
    for (int k = 0; k <= MXB_COL_PROB; ++k)
        if (!host_in->f64[k]) return fail(MXB_EINVAL, "mxb_trace_host: core photon column missing");
    if (chunk <= 0) chunk = 1 << 21;
    if (chunk > n && n > 0) chunk = n;
    memset(status_host, 0, sizeof(unsigned long long) * MXB_STATUS_WORDS);
    if (n == 0) return MXB_OK;

    constexpr int NBUF = 3;
    int nf = 0, ni = 0, nd = 0;
    int fidx[MXB_MAX_F64_COLS], iidx[MXB_MAX_I64_COLS], didx[MXB_MAX_SLOTS];
    for (int k = 0; k < MXB_MAX_F64_COLS; ++k) if (host_in->f64[k] || host_out->f64[k]) fidx[nf++] = k;
    for (int k = 0; k < MXB_MAX_I64_COLS; ++k) if (host_out->i64[k]) iidx[ni++] = k;
    for (int k = 0; k < MXB_MAX_SLOTS; ++k) if (host_in->draws[k]) didx[nd++] = k;
    const size_t planes = (size_t)nf + ni + nd;

    // device staging (program copy, status block, NBUF chunk buffers, streams, events) is cached per
    // host thread and device and only grows: repeated calls do no cudaMalloc / cudaFree
    HostStage& S = host_stage();
    int result = MXB_OK;
#define HTRY(expr)                                                                              \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            result = fail(MXB_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));       \
            goto cleanup;                                                                       \
        }                                                                                       \
    } while (0)
    {
        int dev = 0;
        HTRY(cudaGetDevice(&dev));
        if (S.dev != dev) { S.release(); S.dev = dev; }
        if (!S.s_in) {
            HTRY(cudaStreamCreateWithFlags(&S.s_in, cudaStreamNonBlocking));
            HTRY(cudaStreamCreateWithFlags(&S.s_k, cudaStreamNonBlocking));
            HTRY(cudaStreamCreateWithFlags(&S.s_out, cudaStreamNonBlocking));
            for (int b = 0; b < NBUF; ++b) {
                HTRY(cudaEventCreateWithFlags(&S.ev_in[b], cudaEventDisableTiming));
                HTRY(cudaEventCreateWithFlags(&S.ev_k[b], cudaEventDisableTiming));
                HTRY(cudaEventCreateWithFlags(&S.ev_out[b], cudaEventDisableTiming));
            }
            HTRY(cudaMalloc(&S.dstatus, sizeof(unsigned long long) * MXB_STATUS_WORDS));
        }
        if (S.prog_bytes < prog_words * 8) {
            if (S.dprog) HTRY(cudaFree(S.dprog));
            S.dprog = nullptr;
            HTRY(cudaMalloc(&S.dprog, prog_words * 8));
            S.prog_bytes = prog_words * 8;
        }
        const size_t need = planes * (size_t)chunk * 8;
        if (S.buf_bytes < need) {
            for (int b = 0; b < NBUF; ++b) {
                if (S.dbuf[b]) HTRY(cudaFree(S.dbuf[b]));
                S.dbuf[b] = nullptr;
            }
            S.buf_bytes = 0;
            for (int b = 0; b < NBUF; ++b) HTRY(cudaMalloc(&S.dbuf[b], need));
            S.buf_bytes = need;
        }
        cudaStream_t s_in = S.s_in, s_k = S.s_k, s_out = S.s_out;
        double* dprog = S.dprog;
        unsigned long long* dstatus = S.dstatus;
        HTRY(cudaMemsetAsync(dstatus, 0, sizeof(unsigned long long) * MXB_STATUS_WORDS, s_k));
        HTRY(cudaMemcpyAsync(dprog, prog_host, prog_words * 8, cudaMemcpyHostToDevice, s_k));
        int64_t nchunks = (n + chunk - 1) / chunk;
        for (int64_t c = 0; c < nchunks; ++c) {
            const int b = (int)(c % NBUF);
            const int64_t off = c * chunk;
            const int64_t m = (off + chunk <= n) ? chunk : (n - off);
            char* base = S.dbuf[b];
            MxbColumns dc;
            memset(&dc, 0, sizeof(dc));
            // buffer b is free once its previous D2H finished
            if (c >= NBUF) HTRY(cudaStreamWaitEvent(s_in, S.ev_out[b], 0));
            size_t p = 0;
            for (int k = 0; k < nf; ++k, ++p) {
                dc.f64[fidx[k]] = reinterpret_cast<double*>(base + p * (size_t)chunk * 8);
                if (host_in->f64[fidx[k]])
                    HTRY(cudaMemcpyAsync(dc.f64[fidx[k]], host_in->f64[fidx[k]] + off, (size_t)m * 8,
                                         cudaMemcpyHostToDevice, s_in));
            }
            for (int k = 0; k < ni; ++k, ++p)
                dc.i64[iidx[k]] = reinterpret_cast<long long*>(base + p * (size_t)chunk * 8);
            for (int k = 0; k < nd; ++k, ++p) {
                double* d = reinterpret_cast<double*>(base + p * (size_t)chunk * 8);
                dc.draws[didx[k]] = d;
                HTRY(cudaMemcpyAsync(d, host_in->draws[didx[k]] + off, (size_t)m * 8,
                                     cudaMemcpyHostToDevice, s_in));
            }
            HTRY(cudaEventRecord(S.ev_in[b], s_in));
            HTRY(cudaStreamWaitEvent(s_k, S.ev_in[b], 0));
            rc = launch_trace(dprog, n_ops, stage_words, &dc, m, photon_id0 + off, seed, dstatus, s_k);
            if (rc) { result = rc; goto cleanup; }
            HTRY(cudaEventRecord(S.ev_k[b], s_k));
            HTRY(cudaStreamWaitEvent(s_out, S.ev_k[b], 0));
            for (int k = 0; k < nf; ++k) {
                if (!host_out->f64[fidx[k]]) continue;
                if (fidx[k] == MXB_COL_ENERGY && host_out->f64[fidx[k]] == host_in->f64[fidx[k]])
                    continue;   // energy is never modified: nothing to bring back in place
                HTRY(cudaMemcpyAsync(host_out->f64[fidx[k]] + off, dc.f64[fidx[k]], (size_t)m * 8,
                                     cudaMemcpyDeviceToHost, s_out));
            }
            for (int k = 0; k < ni; ++k)
                HTRY(cudaMemcpyAsync(host_out->i64[iidx[k]] + off, dc.i64[iidx[k]], (size_t)m * 8,
                                     cudaMemcpyDeviceToHost, s_out));
            HTRY(cudaEventRecord(S.ev_out[b], s_out));
        }
        HTRY(cudaStreamSynchronize(s_out));
        HTRY(cudaMemcpyAsync(status_host, dstatus, sizeof(unsigned long long) * MXB_STATUS_WORDS,
                             cudaMemcpyDeviceToHost, s_k));
        HTRY(cudaStreamSynchronize(s_k));
    }
cleanup:
    if (result != MXB_OK) { cudaDeviceSynchronize(); S.release(); }
#undef HTRY
    return result;
}

}  // extern "C"
