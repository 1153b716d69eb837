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
#include "mxb_ops.cuh"
#include "mxb_jit.h"

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
    const double* src[11];   // core planes the photons are READ from (== cols.f64[0..10] in place)
    int store_energy;        // out of place / born photons: energy has to be written to the destination
    int born;                // header flag: the program creates the photons, nothing is read
    MxbColumns cols;
};

// ---------------------------------------------------------------------------
// interpreter-side column stores.  Output columns are resolved once per CTA: cp[k] = device
// pointer of the op's k-th column and a mode word (bit k: materialised, bit 8+k: this op
// initialises the column, i.e. stores NaN / -1 for photons it does not touch - outside arrays,
// or in the first search round).  Per op, `wmask` = the columns this lane has to store.
// ---------------------------------------------------------------------------
struct Ctx {
    const TraceParams* P;
    long long i;           // photon index in the batch
    bool active;
    bool init_round;            // column-initialising stores are live (see put)
    unsigned long long* st_sm;  // smem status accumulators
};

__device__ __forceinline__ double draw(const Ctx& c, int slot, int kind) {
    return draw_value(c.P->cols.draws[slot], c.i, c.P->seed, (unsigned long long)(c.P->id0 + c.i), slot, kind);
}

typedef unsigned long long ColEntry;

__device__ __forceinline__ int store_mask(const Ctx& c, int mode, bool hit) {
    const int sel = hit ? 0xff : (c.init_round ? (mode >> 8) : 0);
    return c.active ? (mode & sel) : 0;
}
__device__ __forceinline__ void put(const Ctx& c, int wmask, int k, ColEntry e, bool hit, double v) {
    if (wmask & (1 << k)) st_global(reinterpret_cast<double*>(e) + c.i, hit ? v : nan64());
}
__device__ __forceinline__ void put_id(const Ctx& c, int wmask, int k, ColEntry e, bool hit, long long v) {
    if (wmask & (1 << k)) st_global(reinterpret_cast<long long*>(e) + c.i, hit ? v : -1LL);
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
    __shared__ unsigned long long hot_keys[MXB_HOT_SLOTS];
    __shared__ double hot_vals[MXB_HOT_SLOTS];
    const HotCache hot{hot_keys, hot_vals};
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
    hot_init(hot, tid, kThreads);
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
        if (ctx.active && !P.born) {
            ph.pos = V3{P.src[0][i], P.src[1][i], P.src[2][i]};
            ph.dir = V3{P.src[3][i], P.src[4][i], P.src[5][i]};
            ph.pol = V3{P.src[6][i], P.src[7][i], P.src[8][i]};
            ph.energy = P.src[9][i];
            ph.prob = P.src[10][i];
        } else {
            ph.pos = ph.dir = ph.pol = V3{kNaN, kNaN, kNaN};
            ph.energy = ph.prob = kNaN;
        }
        photon_loaded(ph);
        ph.ip = V3{kNaN, kNaN, kNaN};
        ph.l0 = ph.l1 = kNaN;

        // array iteration state
        int arr_nhit = 0, arr_pc = -1;
        ArrayIter arr{0, 0, false, false};
        ctx.init_round = true;
        int row = 0;    // word offset of the current facet row (0 = the blob itself outside arrays)
        int geom = 0;   // word offset of the current geometry block

        // leave the array whose ARRAY_BEGIN is op `bpc`
        auto array_exit = [&]() {
            if (arr_nhit >= 2) count_status(st_sm, MXB_ST_MULTI_HIT);
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
                if (ph.hit) {
                    if (op.flags & 1) op_lens_refl(st_sm, ph, pr, P.prog);
                    else op_lens(ph, pr);
                }
                break;
            }
            case MXB_OP_RSCATTER: {
                const OpCold& c = opc[pc];
                const int wm = store_mask(ctx, c.mode, ph.hit);
                double a = 0, b = 0;
                if (ph.hit) {
                    double z0, z1;
                    rscatter_draws(pr[3], pr[4], P.cols.draws[c.s0], P.cols.draws[c.s1], i, P.seed,
                                   (unsigned long long)(P.id0 + i), c.s0, c.s1, z0, z1);
                    op_rscatter(ph, pr, z0, z1, a, b);
                }
                put(ctx, wm, 0, c.cp[0], ph.hit, a);
                put(ctx, wm, 1, c.cp[1], ph.hit, b);
                break;
            }
            case MXB_OP_GSCATTER: {
                const OpCold& c = opc[pc];
                const int wm = store_mask(ctx, c.mode, ph.hit);
                double a = 0;
                if (ph.hit) {
                    const double zn = (op.flags & 2) ? P.cols.f64[c.c[1]][i] : draw(ctx, c.s0, 1);   // callable scatter: angle column
                    const double u = draw(ctx, c.s1, 0);
                    op_gscatter(ph, pr, op.flags, zn, u, a);
                }
                put(ctx, wm, 0, c.cp[0], ph.hit, a);
                break;
            }
            case MXB_OP_FILTER: {
                if (ph.hit) mul_prob(st_sm, ph, filter_value(st_sm, pr, ph.energy, op.flags));
                break;
            }
            case MXB_OP_GFILTER: {
                if (ctx.active) mul_prob(st_sm, ph, filter_value(st_sm, pr, ph.energy, op.flags));
                break;
            }
            case MXB_OP_GRATING: {
                const OpCold& c = opc[pc];
                const int wm = store_mask(ctx, c.mode, ph.hit);
                double order = 0, blaze = 0;
                if (ph.hit) {
                    const double u = draw(ctx, c.s0, 0);
                    const Ref sel = B + op.pg;
                    bool blocked = false;
                    double trans = 0.0;
                    if (op.flags & 8) {   // L1 support: block = openfraction, offset of the Si transmission table
                        const Ref lb = B + c.c[2];
                        blocked = draw(ctx, c.s1, 0) > lb[0];
                        if (blocked) trans = filter_value(st_sm, PRef<false>{P.prog, (int)lb[1]}, ph.energy, 1);
                    }
                    op_grating(st_sm, ph, pr, B + geom, op.flags,
                               [&](double energy, double bl, double& psel) {
                                   return select_order(sel, P.prog, u, energy, bl, psel);
                               },
                               order, blaze, blocked, trans,
                               (op.flags & 16) ? P.cols.f64[c.c[3]][i] : 0.0);   // (input column: plain index)
                }
                put(ctx, wm, 0, c.cp[0], ph.hit, order);
                put(ctx, wm, 1, c.cp[1], ph.hit, blaze);
                break;
            }
            case MXB_OP_DETPIX: {
                // detector.py:73-75; pr: pixsize cp0 cp1; optional fused image: pg: nx ny sel_lo n_sel, s0 image slot
                const OpCold& c = opc[pc];
                const int wm = store_mask(ctx, c.mode, ph.hit);
                double px, py;
                op_detpix(ph, pr, op.flags, px, py);
                put(ctx, wm, 0, c.cp[0], ph.hit, px);
                put(ctx, wm, 1, c.cp[1], ph.hit, py);
                if (c.s0 >= 0 && ph.hit) {
                    const Ref gp = B + op.pg;
                    const long long idn = (op.flags & 1) ? (long long)(B + row)[c.w15] : (long long)c.w14;
                    accumulate_image(hot, P.cols.f64[c.s0], gp, idn, px, py, ph.prob);
                }
                break;
            }
            case MXB_OP_ACIS: {
                // det_acis.py:31-58 ; per-facet pr: pixsize cp0 cp1 sh ct st ox oy ; global: f pixrad odet0 odet1 cosr sinr
                const OpCold& c = opc[pc];
                const int wm = store_mask(ctx, c.mode, ph.hit);
                const Ref gp = B + op.pg;
                double o[8];
                op_acis(ph, pr, gp, o);
                const double chipx = o[0], chipy = o[1];
#pragma unroll
                for (int k = 0; k < 8; ++k) put(ctx, wm, k, c.cp[k], ph.hit, o[k]);
                if (c.s0 >= 0 && ph.hit) {
                    // fused detector image (chip pixel convention is 1-based: det_acis.py:33-34)
                    const long long idn = (long long)(B + row)[c.w15];
                    accumulate_image(hot, P.cols.f64[c.s0], gp + 6, idn, chipx - 1.0, chipy - 1.0, ph.prob);
                }
                break;
            }
            case MXB_OP_GENERATE: {
                const OpCold& c = opc[pc];
                double time = 0, polangle = 0;
                if (ctx.active) {
                    const int pc1 = c.c[1] >= MXB_COL_INIT ? c.c[1] - MXB_COL_INIT : c.c[1];
                    op_generate(ph, pr, P.prog, (unsigned long long)(P.id0 + i), [&](int s) { return draw(ctx, s, 0); },
                                c.s0, c.s1, c.w14, c.w15, time, polangle, op.flags,
                                (op.flags & 1) ? P.cols.f64[MXB_COL_ENERGY][i] : 0.0,
                                ((op.flags & 2) && pc1 >= 0) ? P.cols.f64[pc1][i] : 0.0);
                }
                const int wm = store_mask(ctx, c.mode, true);
                if (!(op.flags & 4)) put(ctx, wm, 0, c.cp[0], true, time);
                if (!(op.flags & 2)) put(ctx, wm, 1, c.cp[1], true, polangle);
                put(ctx, wm, 2, c.cp[2], true, pr[8]);
                put(ctx, wm, 3, c.cp[3], true, pr[9]);
                break;
            }
            case MXB_OP_POINTING: {
                const OpCold& c = opc[pc];
                if (ctx.active) {
                    const MxbColumns& C = P.cols;
                    double ua = 0, zj = 0;
                    if (op.flags & 1) {
                        ua = draw(ctx, c.s0, 0);
                        if (pr[21] > 0.0) zj = draw(ctx, c.s1, 1);
                    }
                    op_pointing(ph, pr, op.flags, C.f64[c.c[0]][i], C.f64[c.c[1]][i], C.f64[c.c[2]][i], ua, zj);
                }
                break;
            }
            case MXB_OP_LABCONE: {
                const OpCold& c = opc[pc];
                if (ctx.active) {
                    const double u0 = draw(ctx, c.s0, 0), u1 = draw(ctx, c.s1, 0);
                    op_labcone(ph, pr, u0, u1, P.cols.f64[c.c[0]][i]);
                }
                break;
            }
            case MXB_OP_FARLAB: {
                const OpCold& c = opc[pc];
                if (ctx.active) {
                    const double u0 = draw(ctx, c.s0, 0), u1 = draw(ctx, c.s1, 0);
                    op_farlab(ph, pr, u0, u1, P.cols.f64[c.c[0]][i]);
                }
                break;
            }
            case MXB_OP_CYLINDER: {
                geom = op.pg;
                ph.hit = cylinder_intersect(B + geom, ph.pos, ph.dir, ph.ip, ph.l0, ph.l1) && ctx.active;
                const unsigned m = __ballot_sync(0xffffffffu, ph.hit);
                if (lane0) my_hits[pc] += __popc(m);
                break;
            }
            case MXB_OP_QFACTOR: {
                if (ph.hit) op_qfactor(st_sm, ph, pr);
                break;
            }
            case MXB_OP_L2ABS: {
                if (ph.hit) op_l2abs(st_sm, ph, pr, B + geom);
                break;
            }
            case MXB_OP_BREWSTER: {
                if (ph.hit) op_brewster(st_sm, ph, pr);
                break;
            }
            case MXB_OP_MLEFF: {
                if (ph.hit) op_mleff(st_sm, ph, pr);
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
                    const double u0 = draw(ctx, c.s0, 0), u1 = draw(ctx, c.s1, 0);
                    op_aperture(st_sm, ph, pr, op.flags, u0, u1);
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
                    array_open(arr, H, B + c.c[6], c.c[0], c.c[3], c.c[4], c.c[5], ph, ctx.active, st_sm);
                }
                // search the next facet (ascending index) this photon hits from its CURRENT state
                const bool found = array_search(arr, B, H, B + c.c[6], B + c.c[7], c.c[2], c.c[1], c.c[0], c.c[4], c.c[5],
                                                ph, row);
                if (found) geom = row;
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
                {
                    const Ref H = B + oph[bpc].pg;      // H[18]: offset of the per-facet limit table, H[19]: certificate mode
                    array_revalidate<true>(arr, H, B + (int)H[18], (int)H[19], ph, arr_nhit, row, c.c[2], c.c[1], c.c[0], st_sm);
                }
                // another search round only if some lane that hit still has facets to test
                ctx.init_round = false;               // later rounds only store for photons that hit
                if (__any_sync(0xffffffffu, ph.hit && arr.cur < arr.end)) {
                    if (!ph.hit) arr.cur = arr.end;   // lanes that found nothing are done with this array
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
            if (P.store_energy) C.f64[9][i] = ph.energy;
        }
    }

    __syncthreads();
    hot_flush(hot, tid, kThreads);
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
polarization_vectors_kernel(const double* dx, const double* dy, const double* dz, const double* angle,
                            double* ox, double* oy, double* oz, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const V3 r = polarization_vector(V3{dx[i], dy[i], dz[i]}, angle[i]);
        ox[i] = r.x;
        oy[i] = r.y;
        oz[i] = r.z;
    }
}

__global__ void __launch_bounds__(256)
hist2d_kernel(const double* x, const double* y, const double* w, const long long* sel, long long sel_lo,
              int n_sel, double x0, double y0, long long n, int nx, int ny, double* img,
              unsigned long long* counts) {
    __shared__ unsigned long long hot_keys[MXB_HOT_SLOTS];
    __shared__ double hot_vals[MXB_HOT_SLOTS];
    const HotCache hot{hot_keys, hot_vals};
    hot_init(hot, threadIdx.x, blockDim.x);
    __syncthreads();
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
        if (img && wv == wv) hot_add(hot, &img[bin], wv, (unsigned)bin);
        if (counts) atomicAdd(&counts[bin], 1ULL);
    }
    __syncthreads();
    hot_flush(hot, threadIdx.x, blockDim.x);
}

// the device random stream, exported: exactly the routines the trace kernels call (device_draw /
// device_draw_normal_pair of mxb_device.cuh, same build flags), so a test or bench.py --verify can feed the
// draws a timed launch used to the CPU oracle.  kind 0 uniform, 1 normal, 2 normal pair (RSCATTER with both
// widths: ONE Philox call yields both deviates)
__global__ void __launch_bounds__(256)
debug_draws_kernel(unsigned long long seed, long long id0, long long n, int slot, int kind, double* out0, double* out1) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long gid = (unsigned long long)(id0 + i);
        if (kind == 2) {
            double z0, z1;
            device_draw_normal_pair(seed, gid, slot, z0, z1);
            out0[i] = z0;
            if (out1) out1[i] = z1;
        } else {
            out0[i] = device_draw(seed, gid, slot, kind);
        }
    }
}

// id columns on the wire as int32 (lean host output: ids are small; -1 stays -1)
__global__ void __launch_bounds__(256)
narrow_i64_kernel(const long long* src, int* dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = (int)src[i];
}

// ---------------------------------------------------------------------------
// event compaction (multi-GPU epilogue): rows with sel >= sel_min (and weight > 0) of up to 64
// 8-byte planes are packed densely, order preserved.  Three passes: per-block counts, one-block scan,
// scatter (flags are recomputed: cheaper than storing them).  HBM-bound: 8 B read per row for the
// selector (+8 weight) twice, 8 B read + 8 B written per selected row and plane.
// ---------------------------------------------------------------------------
constexpr int kCompactThreads = 1024;
constexpr int kCompactMaxPlanes = 64;
struct CompactPlanes {
    const unsigned long long* src[kCompactMaxPlanes];
    unsigned long long* dst[kCompactMaxPlanes];
};
__device__ __forceinline__ bool compact_keep(const long long* sel, long long sel_min, const double* w, long long i,
                                             long long n) {
    if (i >= n) return false;
    bool k = sel ? (sel[i] >= sel_min) : true;
    if (w) k = k && (w[i] > 0.0);   // NaN weights are dropped
    return k;
}
__global__ void __launch_bounds__(kCompactThreads)
compact_count_kernel(const long long* sel, long long sel_min, const double* w, long long n, long long* block_counts) {
    const long long i = (long long)blockIdx.x * kCompactThreads + threadIdx.x;
    const int c = __syncthreads_count(compact_keep(sel, sel_min, w, i, n) ? 1 : 0);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}
// cursor != nullptr (append mode): the exclusive offsets start at cursor[0], which then advances by the
// number of rows kept; rows that would land at or beyond `capacity` are dropped by the scatter kernel and
// counted in cursor[1]
__global__ void __launch_bounds__(1024)
compact_scan_kernel(long long* block_counts, long long n_blocks, long long* n_out, long long* cursor = nullptr,
                    long long capacity = 0) {
    // exclusive scan in place, one CTA: chunks of 1024 with a running carry
    __shared__ long long warp_sum[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = cursor ? cursor[0] : 0;
    __syncthreads();
    const long long start = carry;
    for (long long base = 0; base < n_blocks; base += 1024) {
        const long long k = base + threadIdx.x;
        const long long v = k < n_blocks ? block_counts[k] : 0;
        long long x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, x, d);
            if ((threadIdx.x & 31) >= d) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            long long ws = warp_sum[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long y = __shfl_up_sync(0xffffffffu, ws, d);
                if (threadIdx.x >= d) ws += y;
            }
            warp_sum[threadIdx.x] = ws;   // inclusive over warps
        }
        __syncthreads();
        const long long before = ((threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0) + carry;
        if (k < n_blocks) block_counts[k] = before + x - v;   // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (n_out) *n_out = carry - start;
        if (cursor) {
            const long long end = carry > capacity ? capacity : carry;
            if (carry > capacity) cursor[1] += carry - (start > capacity ? start : capacity);
            cursor[0] = end;
        }
    }
}
__global__ void __launch_bounds__(kCompactThreads)
compact_scatter_kernel(const __grid_constant__ CompactPlanes P, int n_planes, const long long* sel, long long sel_min,
                       const double* w, long long n, const long long* block_offsets,
                       long long capacity = 0x7fffffffffffffffLL) {
    __shared__ int warp_cnt[kCompactThreads / 32];
    const long long i = (long long)blockIdx.x * kCompactThreads + threadIdx.x;
    const bool keep = compact_keep(sel, sel_min, w, i, n);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
        int c = warp_cnt[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, c, d);
            if (lane >= d) c += y;
        }
        warp_cnt[lane] = c;   // inclusive
    }
    __syncthreads();
    if (!keep) return;
    const long long o = block_offsets[blockIdx.x] + (warp ? warp_cnt[warp - 1] : 0) + __popc(m & ((1u << lane) - 1u));
    if (o >= capacity) return;
    for (int p = 0; p < n_planes; ++p) P.dst[p][o] = P.src[p][i];
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

thread_local std::string g_kernel_info = "none";

int launch_interp(const double* prog_dev, int n_ops, int stage_words, bool born, const double* const* src,
                  const MxbColumns* cols, int64_t n, int64_t id0, uint64_t seed, unsigned long long* status_dev,
                  cudaStream_t stream) {
    TraceParams P;
    P.prog = prog_dev;
    P.n_ops = n_ops;
    P.stage_words = stage_words;
    P.n = n;
    P.id0 = id0;
    P.seed = seed;
    P.status = status_dev;
    P.cols = *cols;
    P.store_energy = 0;
    for (int k = 0; k <= MXB_COL_PROB; ++k) {
        P.src[k] = src ? src[k] : cols->f64[k];
        if (k == MXB_COL_ENERGY && P.src[k] != cols->f64[k]) P.store_energy = 1;
    }
    P.born = born ? 1 : 0;
    if (born) P.store_energy = 1;
    const size_t stage_bytes = (size_t)stage_words * 8;
    const int grid = grid_for(n, kThreads, 1);
    if (stage_bytes <= (size_t)kMaxSmemStageBytes) {
        // the attribute is per device (and the call is cheap next to a launch): set it whenever this
        // thread's current device changes, so one process can drive several GPUs
        static thread_local int attr_dev = -1;
        int dev = -1;
        CUDA_TRY(cudaGetDevice(&dev));
        if (attr_dev != dev) {
            CUDA_TRY(cudaFuncSetAttribute(mxb_trace_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          kMaxSmemStageBytes));
            attr_dev = dev;
        }
        mxb_trace_kernel<true><<<grid, kThreads, stage_bytes, stream>>>(P);
    } else {
        mxb_trace_kernel<false><<<grid, kThreads, 0, stream>>>(P);
    }
    CUDA_TRY(cudaGetLastError());
    g_kernel_info = "interpreter";
    return MXB_OK;
}

#ifdef MXB_FAST
constexpr bool kFastBuild = true;
#else
constexpr bool kFastBuild = false;
#endif

// kernel selection (see mxb_set_jit in mxb.h): specialised kernel or the interpreter
int launch_trace(const double* prog_dev, const double* prog_host, size_t prog_words, int n_ops, int stage_words,
                 const double* const* src, const MxbColumns* cols, int64_t n, int64_t id0, uint64_t seed,
                 unsigned long long* status_dev, cudaStream_t stream) {
    const mxbjit::Mode mode = mxbjit::mode();
    if (mode == mxbjit::kForce || (mode == mxbjit::kAuto && n >= mxbjit::auto_threshold())) {
        std::string err;
        bool unavailable = false;
        const int rc = mxbjit::launch(prog_dev, prog_host, prog_words, n_ops, stage_words, src, cols, n, id0, seed,
                                      status_dev, stream, kFastBuild, &err, &unavailable);
        if (rc == MXB_OK) {
            g_kernel_info = mxbjit::last_info();
            return MXB_OK;
        }
        if (mode == mxbjit::kForce || !unavailable) return fail(rc, err);
        // auto mode without NVRTC on this machine: the interpreter kernel runs the program
    }
    return launch_interp(prog_dev, n_ops, stage_words, ((int)prog_host[5] & 1) != 0, src, cols, n, id0, seed, status_dev,
                         stream);
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

void mxb_set_jit(int mode) { mxbjit::set_mode(mode); }
int mxb_get_jit(void) { return (int)mxbjit::mode(); }
const char* mxb_jit_info(void) { return g_kernel_info.c_str(); }

long long mxb_jit_source(const double* prog_host, size_t prog_words, const MxbColumns* cols, char* buf,
                         size_t buf_len) {
    int n_ops = 0, stage_words = 0;
    const int rc = validate_program(prog_host, prog_words, &n_ops, &stage_words);
    if (rc) return rc;
    if (!cols) return fail(MXB_EINVAL, "mxb_jit_source: null pointer");
    std::string err;
    const std::string src = mxbjit::source_for(prog_host, prog_words, cols, &err);
    if (src.empty()) return fail(MXB_EJIT, err);
    if (buf && buf_len) {
        const size_t m = src.size() < buf_len - 1 ? src.size() : buf_len - 1;
        memcpy(buf, src.data(), m);
        buf[m] = 0;
    }
    return (long long)src.size();
}

long long mxb_jit_compile(const double* prog_host, size_t prog_words, const MxbColumns* cols) {
    int n_ops = 0, stage_words = 0;
    const int rc = validate_program(prog_host, prog_words, &n_ops, &stage_words);
    if (rc) return rc;
    if (!cols) return fail(MXB_EINVAL, "mxb_jit_compile: null pointer");
    std::string err, info;
    const long long sz = mxbjit::compile_only(prog_host, prog_words, cols, kFastBuild, &info, &err);
    if (sz < 0) return fail((int)sz, err);
    g_kernel_info = info;
    return sz;
}

size_t mxb_compact_workspace(int64_t n) {
    const int64_t blocks = (n + kCompactThreads - 1) / kCompactThreads;
    return (size_t)(blocks + 1) * sizeof(long long);
}

int mxb_compact_events(const void* const* src_planes, void* const* dst_planes, int n_planes, const long long* sel,
                       long long sel_min, const double* weight, int64_t n, long long* n_out_dev, void* workspace,
                       size_t workspace_bytes, void* stream) {
    if (!src_planes || !dst_planes || !n_out_dev || !workspace) return fail(MXB_EINVAL, "mxb_compact_events: null pointer");
    if (n_planes < 1 || n_planes > kCompactMaxPlanes) return fail(MXB_EINVAL, "mxb_compact_events: 1..64 planes");
    if (n < 0) return fail(MXB_EINVAL, "negative n");
    if (workspace_bytes < mxb_compact_workspace(n)) return fail(MXB_EINVAL, "mxb_compact_events: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        CUDA_TRY(cudaMemsetAsync(n_out_dev, 0, sizeof(long long), s));
        return MXB_OK;
    }
    CompactPlanes P;
    memset(&P, 0, sizeof(P));
    for (int p = 0; p < n_planes; ++p) {
        if (!src_planes[p] || !dst_planes[p]) return fail(MXB_EINVAL, "mxb_compact_events: null plane");
        P.src[p] = static_cast<const unsigned long long*>(src_planes[p]);
        P.dst[p] = static_cast<unsigned long long*>(dst_planes[p]);
    }
    const long long blocks = (n + kCompactThreads - 1) / kCompactThreads;
    long long* counts = static_cast<long long*>(workspace);
    compact_count_kernel<<<(unsigned)blocks, kCompactThreads, 0, s>>>(sel, sel_min, weight, n, counts);
    compact_scan_kernel<<<1, 1024, 0, s>>>(counts, blocks, n_out_dev);
    compact_scatter_kernel<<<(unsigned)blocks, kCompactThreads, 0, s>>>(P, n_planes, sel, sel_min, weight, n, counts);
    CUDA_TRY(cudaGetLastError());
    return MXB_OK;
}

int mxb_compact_append(const void* const* src_planes, void* const* dst_planes, int n_planes, const long long* sel,
                       long long sel_min, const double* weight, int64_t n, long long* cursor_dev, int64_t capacity,
                       void* workspace, size_t workspace_bytes, void* stream) {
    if (!src_planes || !dst_planes || !cursor_dev || !workspace) return fail(MXB_EINVAL, "mxb_compact_append: null pointer");
    if (n_planes < 1 || n_planes > kCompactMaxPlanes) return fail(MXB_EINVAL, "mxb_compact_append: 1..64 planes");
    if (n < 0 || capacity < 0) return fail(MXB_EINVAL, "negative n or capacity");
    if (workspace_bytes < mxb_compact_workspace(n)) return fail(MXB_EINVAL, "mxb_compact_append: workspace too small");
    if (n == 0) return MXB_OK;
    CompactPlanes P;
    memset(&P, 0, sizeof(P));
    for (int p = 0; p < n_planes; ++p) {
        if (!src_planes[p] || !dst_planes[p]) return fail(MXB_EINVAL, "mxb_compact_append: null plane");
        P.src[p] = static_cast<const unsigned long long*>(src_planes[p]);
        P.dst[p] = static_cast<unsigned long long*>(dst_planes[p]);
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long long blocks = (n + kCompactThreads - 1) / kCompactThreads;
    long long* counts = static_cast<long long*>(workspace);
    compact_count_kernel<<<(unsigned)blocks, kCompactThreads, 0, s>>>(sel, sel_min, weight, n, counts);
    compact_scan_kernel<<<1, 1024, 0, s>>>(counts, blocks, nullptr, cursor_dev, capacity);
    compact_scatter_kernel<<<(unsigned)blocks, kCompactThreads, 0, s>>>(P, n_planes, sel, sel_min, weight, n, counts, capacity);
    CUDA_TRY(cudaGetLastError());
    return MXB_OK;
}

int mxb_debug_draws(uint64_t seed, int64_t photon_id0, int64_t n, int slot, int kind, double* out0, double* out1,
                    void* stream) {
    if (!out0 || kind < 0 || kind > 2 || slot < 0 || slot >= MXB_MAX_SLOTS) return fail(MXB_EINVAL, "mxb_debug_draws: bad argument");
    if (n <= 0) return n == 0 ? MXB_OK : fail(MXB_EINVAL, "negative n");
    debug_draws_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(seed, photon_id0, n, slot, kind, out0, out1);
    CUDA_TRY(cudaGetLastError());
    return MXB_OK;
}

// device math routines of this build, exported for verification (tests): out[i] = f(x[i], y[i])
__global__ void debug_math_kernel(int kind, const double* x, const double* y, double* out, int64_t n, double log_hi, double log_lo) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double a = x[i], b = y ? y[i] : 0.0;
        double r;
        switch (kind) {
        case 0: r = mxb::m_acos(a); break;
        case 1: r = mxb::m_asin(a); break;
        case 2: r = mxb::pow_loghost(a, log_hi, log_lo, b); break;
        case 3: r = mxb::div(a, b); break;
        case 4: case 5: { double s, c; mxb::sincos_small(a, &s, &c); r = (kind == 4) ? s : c; break; }
        case 6: case 7: { double s, c; mxb::sincos_turn(a, &s, &c); r = (kind == 6) ? s : c; break; }
        default: { double s, c; mxb::m_sincos(a, &s, &c); r = (kind == 8) ? s : c; }
        }
        out[i] = r;
    }
}

int mxb_debug_math(int kind, const double* x, const double* y, double* out, int64_t n, double log_hi, double log_lo, void* stream) {
    if (!x || !out || kind < 0 || kind > 9 || ((kind == 2 || kind == 3) && !y)) return fail(MXB_EINVAL, "mxb_debug_math: bad argument");
    if (n <= 0) return n == 0 ? MXB_OK : fail(MXB_EINVAL, "negative n");
    debug_math_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(kind, x, y, out, n, log_hi, log_lo);
    CUDA_TRY(cudaGetLastError());
    return MXB_OK;
}

int mxb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mxb_trace_from(const double* prog_dev, size_t prog_words, const double* prog_host, const double* const* src_core,
                   const MxbColumns* cols, int64_t n, int64_t photon_id0, uint64_t seed,
                   unsigned long long* status_dev, void* stream) {
    int n_ops = 0, stage_words = 0;
    const int rc = validate_program(prog_host, prog_words, &n_ops, &stage_words);
    if (rc) return rc;
    if (!prog_dev || !cols || !status_dev) return fail(MXB_EINVAL, "mxb_trace: null pointer");
    if (n < 0) return fail(MXB_EINVAL, "mxb_trace: negative photon count");
    if (n == 0) return MXB_OK;
    for (int k = 0; k <= MXB_COL_PROB; ++k) {
        if (!cols->f64[k]) return fail(MXB_EINVAL, "mxb_trace: core photon column missing");
        if (src_core && !src_core[k]) return fail(MXB_EINVAL, "mxb_trace_from: source plane missing");
    }
    return launch_trace(prog_dev, prog_host, prog_words, n_ops, stage_words, src_core, cols, n, photon_id0, seed,
                        status_dev, (cudaStream_t)stream);
}

int mxb_trace(const double* prog_dev, size_t prog_words, const double* prog_host, const MxbColumns* cols,
              int64_t n, int64_t photon_id0, uint64_t seed, unsigned long long* status_dev, void* stream) {
    return mxb_trace_from(prog_dev, prog_words, prog_host, nullptr, cols, n, photon_id0, seed, status_dev, stream);
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

int mxb_polarization_vectors(const double* const dir[3], const double* angle, double* const pol[3], int64_t n,
                             void* stream) {
    if (!dir || !angle || !pol) return fail(MXB_EINVAL, "null pointer");
    if (n <= 0) return n == 0 ? MXB_OK : fail(MXB_EINVAL, "negative n");
    polarization_vectors_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(dir[0], dir[1], dir[2], angle,
                                                                                      pol[0], pol[1], pol[2], n);
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
    return mxb_trace_host_opts(prog_host, prog_words, host_in, host_out, n, chunk, photon_id0, seed, status_host, nullptr);
}

int mxb_trace_host_opts(const double* prog_host, size_t prog_words, const MxbColumns* host_in,
                        const MxbColumns* host_out, int64_t n, int64_t chunk, int64_t photon_id0,
                        uint64_t seed, unsigned long long* status_host, const MxbHostOptions* opt) {
    int n_ops = 0, stage_words = 0;
    int rc = validate_program(prog_host, prog_words, &n_ops, &stage_words);
    if (rc) return rc;
    if (!host_in || !host_out || !status_host) return fail(MXB_EINVAL, "mxb_trace_host: null pointer");
    if (n < 0) return fail(MXB_EINVAL, "negative n");
    for (int k = 0; k <= MXB_COL_PROB; ++k)
        if (!host_in->f64[k]) return fail(MXB_EINVAL, "mxb_trace_host: core photon column missing");
    if (chunk <= 0) chunk = 1 << 20;   // 1 Mi photons: short pipeline fill, copies still >= 8 MB each
    if (chunk > n && n > 0) chunk = n;
    memset(status_host, 0, sizeof(unsigned long long) * MXB_STATUS_WORDS);
    if (n == 0) return MXB_OK;

    constexpr int NBUF = 3;
    int nf = 0, ni = 0, nd = 0;
    int fidx[MXB_MAX_F64_COLS], iidx[MXB_MAX_I64_COLS], didx[MXB_MAX_SLOTS];
    for (int k = 0; k < MXB_MAX_F64_COLS; ++k) if (host_in->f64[k] || host_out->f64[k]) fidx[nf++] = k;
    int n32 = 0;        // id columns that travel as int32 (opt->i32_out): one more half-size plane each
    for (int k = 0; k < MXB_MAX_I64_COLS; ++k) {
        const bool narrow = opt && opt->i32_out[k];
        if (host_out->i64[k] || narrow) iidx[ni++] = k;
        if (narrow) ++n32;
    }
    for (int k = 0; k < MXB_MAX_SLOTS; ++k) if (host_in->draws[k]) didx[nd++] = k;
    const size_t planes = (size_t)nf + ni + nd + (size_t)(n32 + 1) / 2;

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
            rc = launch_trace(dprog, prog_host, prog_words, n_ops, stage_words, nullptr, &dc, m, photon_id0 + off,
                              seed, dstatus, s_k);
            if (rc) { result = rc; goto cleanup; }
            int* narrow_base = reinterpret_cast<int*>(base + p * (size_t)chunk * 8);   // after the 8-byte planes
            {
                int j = 0;
                for (int k = 0; k < ni; ++k) {
                    if (!(opt && opt->i32_out[iidx[k]])) continue;
                    narrow_i64_kernel<<<grid_for(m, 256, 8), 256, 0, s_k>>>(dc.i64[iidx[k]], narrow_base + (size_t)j * chunk, m);
                    ++j;
                }
                HTRY(cudaGetLastError());
            }
            HTRY(cudaEventRecord(S.ev_k[b], s_k));
            HTRY(cudaStreamWaitEvent(s_out, S.ev_k[b], 0));
            for (int k = 0; k < nf; ++k) {
                if (!host_out->f64[fidx[k]]) continue;
                if (fidx[k] == MXB_COL_ENERGY && host_out->f64[fidx[k]] == host_in->f64[fidx[k]])
                    continue;   // energy is never modified: nothing to bring back in place
                HTRY(cudaMemcpyAsync(host_out->f64[fidx[k]] + off, dc.f64[fidx[k]], (size_t)m * 8,
                                     cudaMemcpyDeviceToHost, s_out));
            }
            {
                int j = 0;
                for (int k = 0; k < ni; ++k) {
                    if (opt && opt->i32_out[iidx[k]]) {
                        HTRY(cudaMemcpyAsync(opt->i32_out[iidx[k]] + off, narrow_base + (size_t)j * chunk, (size_t)m * 4,
                                             cudaMemcpyDeviceToHost, s_out));
                        ++j;
                    }
                    if (host_out->i64[iidx[k]])
                        HTRY(cudaMemcpyAsync(host_out->i64[iidx[k]] + off, dc.i64[iidx[k]], (size_t)m * 8,
                                             cudaMemcpyDeviceToHost, s_out));
                }
            }
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
