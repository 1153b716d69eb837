// mxb_jit.cpp — program -> specialised sm_100a kernel (NVRTC), cache, launch.  See mxb_jit.h.
#include "mxb_jit.h"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <unordered_map>
#include <vector>

#define MXB_PIPE_BYTES_PER_WARP (11 * 32 * 8)   // mxb_ops.cuh: MXB_PIPE_WORDS_PER_WARP doubles

#include "mxb_embed.inc"   // kSrcMxbH, kSrcDeviceCuh, kSrcOpsCuh : the headers, embedded at build time

namespace mxbjit {
namespace {

// ---------------------------------------------------------------------------
// NVRTC through dlopen: libmxb has no link-time dependency on it
// ---------------------------------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
    void* h = nullptr;
    std::string path, why;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    int (*DestroyProgram)(nvrtcProgram*) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*Version)(int*, int*) = nullptr;
    int major = 0, minor = 0;
};

Nvrtc& nvrtc() {
    static Nvrtc N;
    static std::once_flag once;
    std::call_once(once, [] {
        std::vector<std::string> cand;
        if (const char* e = getenv("MXB_NVRTC_PATH")) cand.push_back(e);
        cand.push_back("libnvrtc.so.12");
        cand.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
        cand.push_back("/usr/local/cuda/lib64/libnvrtc.so");
        cand.push_back("libnvrtc.so");
        for (const auto& c : cand) {
            N.h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
            if (N.h) { N.path = c; break; }
        }
        if (!N.h) { N.why = "libnvrtc.so.12 not found (set MXB_NVRTC_PATH)"; return; }
#define SYM(name)                                                                      \
        *(void**)(&N.name) = dlsym(N.h, "nvrtc" #name);                                \
        if (!N.name) { N.why = "nvrtc" #name " missing in " + N.path; N.h = nullptr; return; }
        SYM(CreateProgram) SYM(CompileProgram) SYM(DestroyProgram) SYM(GetCUBINSize) SYM(GetCUBIN)
        SYM(GetProgramLogSize) SYM(GetProgramLog) SYM(GetErrorString) SYM(Version)
#undef SYM
        N.Version(&N.major, &N.minor);
    });
    return N;
}

uint64_t fnv1a(const void* p, size_t n, uint64_t h = 1469598103934665603ULL) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t k = 0; k < n; ++k) { h ^= b[k]; h *= 1099511628211ULL; }
    return h;
}

// ---------------------------------------------------------------------------
// code generator
// ---------------------------------------------------------------------------
struct Op {
    int type, flags, pg, pf, c[8], s0, s1, w14, w15;
};

struct Gen {
    const double* W;
    size_t words;
    int n_ops, stage_words;
    const MxbColumns* cols;
    bool emit;
    bool staged;                  // facet rows / tables come from shared memory (else global)
    std::string src, key, err;
    std::vector<int> fmap, imap, dmap, smap;
    std::map<int, int> fidx, iidx, didx;
    std::map<std::pair<int, int>, int> sidx;
    bool need_blob = false;       // some op reads the blob through PRef (staging needed)
    bool need_hot = false;        // a detector image is accumulated: per-CTA hot-pixel cache
    bool born = false;            // header flag: the program creates the photons (no loads)
    int threads = 640;            // CTA size the kernel is compiled for
    bool pipe_wanted = true, pipe = false;
    bool idx64 = false;      // photon indices as long long (launches of >= 2^31 photons)
    std::vector<Op> ops;
    // array context
    bool in_array = false;
    std::string geom;             // accessor expression of the current geometry block
    // stores already emitted since the last hit-defining op (layers of a FlatStack repeat the same
    // loc-coos / id / pos commit: optics/base.py:261-281)
    std::set<std::pair<int, std::string>> stored;
    bool pos_committed = false;
    void new_hit() { stored.clear(); pos_committed = false; }
    bool repeat(int key, const std::string& val) {
        if (val != "ph.l0" && val != "ph.l1" && val.find("LL") == std::string::npos && val.find("(long long)") != 0) return false;
        return !stored.insert({key, val}).second;
    }
    void commit_pos() {
        if (pos_committed) return;
        pos_committed = true;
        out("            if (ph.hit) ph.pos = ph.ip;");
    }

    void out(const char* fmt, ...) {
        if (!emit) return;
        char buf[2048];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        src += buf;
        src += '\n';
    }
    void keyi(long long v) { key.append(reinterpret_cast<const char*>(&v), sizeof(v)); }
    bool fail(const std::string& m) { if (err.empty()) err = m; return false; }
    bool range_ok(long long off, long long count) {
        return off >= 0 && count >= 0 && (size_t)(off + count) <= words;
    }
    // a blob word whose VALUE is compiled into the kernel (sizes, kinds): part of the cache key
    int baked(int off) {
        if (!range_ok(off, 1)) { fail("program: offset outside the blob"); return 0; }
        const int v = (int)W[off];
        keyi(v);
        return v;
    }
    // ---- parameter accessors ----
    std::string S(int off, int count) {   // kernel-parameter scalars (constant bank)
        if (!range_ok(off, count)) { fail("program: parameter block outside the blob"); return "SRef{P.s}"; }
        auto it = sidx.find({off, count});
        int base;
        if (it != sidx.end()) base = it->second;
        else {
            base = (int)smap.size();
            for (int k = 0; k < count; ++k) smap.push_back(off + k);
            sidx[{off, count}] = base;
        }
        return "SRef{P.s + " + std::to_string(base) + "}";
    }
    std::string Bref(int off) {           // staged blob, absolute word offset
        need_blob = true;
        if (!range_ok(off, 1)) fail("program: table offset outside the blob");
        return "(B + " + std::to_string(off) + ")";
    }
    std::string Rref(int pf) {            // inside the current facet row
        need_blob = true;
        return "(B + (row + " + std::to_string(pf) + "))";
    }
    // op parameters: per-facet (row-relative) inside arrays, scalars otherwise
    std::string PR(const Op& o, int count) {
        if (o.pf >= 0) return in_array ? Rref(o.pf) : S(o.pf, count);
        return S(o.pg >= 0 ? o.pg : 0, count);
    }
    int pr_off(const Op& o) { return o.pf >= 0 ? o.pf : (o.pg >= 0 ? o.pg : 0); }

    // ---- columns ----
    bool has_f(int col) {
        if (col < 0) return false;
        const int idx = col >= MXB_COL_INIT ? col - MXB_COL_INIT : col;
        return idx < MXB_MAX_F64_COLS && cols->f64[idx] != nullptr;
    }
    bool has_i(int col) {
        if (col < 0) return false;
        const int idx = col >= MXB_COL_INIT ? col - MXB_COL_INIT : col;
        return idx < MXB_MAX_I64_COLS && cols->i64[idx] != nullptr;
    }
    std::string F(int idx) {
        auto it = fidx.find(idx);
        if (it == fidx.end()) { fidx[idx] = (int)fmap.size(); fmap.push_back(idx); it = fidx.find(idx); }
        return "P.f[" + std::to_string(it->second) + "]";
    }
    std::string I(int idx) {
        auto it = iidx.find(idx);
        if (it == iidx.end()) { iidx[idx] = (int)imap.size(); imap.push_back(idx); it = iidx.find(idx); }
        return "P.i[" + std::to_string(it->second) + "]";
    }
    std::string D(int slot) {             // injected draw array of a slot, or a literal null -> Philox
        if (slot < 0 || slot >= MXB_MAX_SLOTS || !cols->draws[slot]) return "((const double*)0)";
        auto it = didx.find(slot);
        if (it == didx.end()) { didx[slot] = (int)dmap.size(); dmap.push_back(slot); it = didx.find(slot); }
        return "P.d[" + std::to_string(it->second) + "]";
    }
    bool injected(int slot) { return slot >= 0 && slot < MXB_MAX_SLOTS && cols->draws[slot]; }
    std::string draw(int slot, int kind) {
        return "draw_value(" + D(slot) + ", i, P.seed, gid, " + std::to_string(slot) + ", " + std::to_string(kind) + ")";
    }
    // store condition of a column reference (see mxb_trace.cu store_mask)
    std::string cond(int col) {
        if (col >= MXB_COL_INIT) return in_array ? "(ph.hit || (init_round && active))" : "active";
        return "ph.hit";
    }
    void put(int col, const std::string& val) {
        if (!has_f(col)) return;
        const int idx = col >= MXB_COL_INIT ? col - MXB_COL_INIT : col;
        if (repeat(idx, val)) return;
        out("            jput(%s, i, %s, ph.hit, %s);", F(idx).c_str(), cond(col).c_str(), val.c_str());
    }
    void put_id(int col, const std::string& val) {
        if (!has_i(col)) return;
        const int idx = col >= MXB_COL_INIT ? col - MXB_COL_INIT : col;
        if (repeat(1000 + idx, val)) return;
        out("            jput_id(%s, i, %s, ph.hit, %s);", I(idx).c_str(), cond(col).c_str(), val.c_str());
    }
    std::string row_id(const Op& o, bool from_row) {
        if (from_row) return "(long long)(B + row)[" + std::to_string(o.w15) + "]";
        return std::to_string(o.w14) + "LL";
    }
    void hit_count(int pc) { out("            h%d += __popc(__ballot_sync(0xffffffffu, ph.hit));", pc); }

    bool parse() {
        if (!W || words < MXB_HEADER_WORDS) return fail("program: missing header");
        if ((size_t)(MXB_HEADER_WORDS + n_ops * MXB_OP_WORDS) > words) return fail("program: truncated op table");
        ops.resize(n_ops);
        for (int k = 0; k < n_ops; ++k) {
            const double* w = W + MXB_HEADER_WORDS + k * MXB_OP_WORDS;
            Op& o = ops[k];
            o.type = (int)w[0]; o.flags = (int)w[1]; o.pg = (int)w[2]; o.pf = (int)w[3];
            for (int j = 0; j < 8; ++j) o.c[j] = (int)w[4 + j];
            o.s0 = (int)w[12]; o.s1 = (int)w[13]; o.w14 = (int)w[14]; o.w15 = (int)w[15];
        }
        // structure key: the op table, which pointers exist, staging
        key.assign(reinterpret_cast<const char*>(W + MXB_HEADER_WORDS), (size_t)n_ops * MXB_OP_WORDS * 8);
        for (int k = 0; k < MXB_MAX_F64_COLS; ++k) key.push_back(cols->f64[k] ? 1 : 0);
        for (int k = 0; k < MXB_MAX_I64_COLS; ++k) key.push_back(cols->i64[k] ? 1 : 0);
        for (int k = 0; k < MXB_MAX_SLOTS; ++k) key.push_back(cols->draws[k] ? 1 : 0);
        keyi(stage_words);
        born = ((int)W[5] & 1) != 0;
        keyi(born ? 1 : 0);
        return true;
    }

    void folded_commit(const Op& o) {
        if (!(o.flags & 256)) return;
        // optics/base.py:201-209 folded into the element's op: c5,c6 loc-coos columns, c7 id column
        put(o.c[5], "ph.l0");
        put(o.c[6], "ph.l1");
        put_id(o.c[7], row_id(o, (o.flags & 512) != 0));
        commit_pos();
    }

    bool gen_op(int pc) {
        const Op& o = ops[pc];
        const int fl = o.flags & 255;
        out("            // ---- op %d: type %d flags %d", pc, o.type, o.flags);
        switch (o.type) {
        case MXB_OP_PLANE: {
            if (in_array) return fail("PLANE inside an array");
            geom = S(o.pg, 14);
            new_hit();
            out("            ph.hit = plane_intersect(%s, ph.pos, ph.dir, %s, ph.ip, ph.l0, ph.l1) && active;",
                geom.c_str(), (fl & 1) ? "true" : "false");
            hit_count(pc);
            break;
        }
        case MXB_OP_LOADHIT: {
            for (int j = 0; j < 6; ++j)
                if (o.c[j] < 0 || o.c[j] >= MXB_MAX_F64_COLS || !cols->f64[o.c[j]]) return fail("LOADHIT: missing input column");
            geom = S(o.pg, 14);
            new_hit();
            out("            ph.hit = false;");
            out("            if (active) {");
            out("                ph.hit = %s[i] != 0.0;", F(o.c[0]).c_str());
            out("                ph.ip = V3{%s[i], %s[i], %s[i]};", F(o.c[1]).c_str(), F(o.c[2]).c_str(), F(o.c[3]).c_str());
            out("                ph.l0 = %s[i];", F(o.c[4]).c_str());
            out("                ph.l1 = %s[i];", F(o.c[5]).c_str());
            out("            }");
            hit_count(pc);
            break;
        }
        case MXB_OP_COMMIT: {
            put(o.c[0], "ph.l0");
            put(o.c[1], "ph.l1");
            put_id(o.c[2], row_id(o, (fl & 1) != 0));
            commit_pos();
            break;
        }
        case MXB_OP_BAFFLE:
            out("            if (ph.hit) ph.pos = ph.ip; else ph.prob = 0.0;");
            break;
        case MXB_OP_LENS:
            if (o.flags & 1) out("            if (ph.hit) op_lens_refl(st_sm, ph, %s, P.prog);", PR(o, 7).c_str());
            else out("            if (ph.hit) op_lens(ph, %s);", PR(o, 4).c_str());
            break;
        case MXB_OP_RSCATTER: {
            const std::string p = PR(o, 5);
            // whether a width is zero is part of the structure (a zero width skips its rotation and its draw):
            // known at compile time for a single element, tested per photon for per-facet parameter rows
            int nz_in = -1, nz_perp = -1;
            if (!(in_array && o.pf >= 0) && range_ok(pr_off(o), 5)) {
                nz_in = W[pr_off(o) + 3] != 0.0 ? 1 : 0;
                nz_perp = W[pr_off(o) + 4] != 0.0 ? 1 : 0;
                keyi(nz_in);
                keyi(nz_perp);
            }
            out("            {");
            out("            double a = 0, b = 0;");
            out("            if (ph.hit) {");
            out("                double z0, z1;");
            out("                rscatter_draws<%d, %d>(%s[3], %s[4], %s, %s, i, P.seed, gid, %d, %d, z0, z1);", nz_in, nz_perp, p.c_str(),
                p.c_str(), D(o.s0).c_str(), D(o.s1).c_str(), o.s0, o.s1);
            out("                op_rscatter<%d, %d>(ph, %s, z0, z1, a, b);", nz_in, nz_perp, p.c_str());
            out("            }");
            put(o.c[0], "a");
            put(o.c[1], "b");
            out("            }");
            break;
        }
        case MXB_OP_GSCATTER: {
            out("            {");
            out("            double a = 0;");
            out("            if (ph.hit) {");
            if (fl & 2) out("                const double zn = %s[i];   // callable scatter: angle column", F(o.c[1]).c_str());
            else out("                const double zn = %s;", draw(o.s0, 1).c_str());
            out("                const double u = %s;", draw(o.s1, 0).c_str());
            out("                op_gscatter(ph, %s, %d, zn, u, a);", PR(o, 1).c_str(), fl);
            out("            }");
            put(o.c[0], "a");
            out("            }");
            break;
        }
        case MXB_OP_FILTER:
        case MXB_OP_GFILTER: {
            const int off = pr_off(o);
            if (in_array && o.pf >= 0) return fail("FILTER inside an array row");
            const int n = baked(off);
            const std::string p = (n == 0) ? S(off, 2) : Bref(off);
            if (n != 0 && !range_ok(off, 1 + 2 * (long long)n)) return fail("FILTER: table outside the blob");
            out("            if (%s) mul_prob(st_sm, ph, filter_value(st_sm, %s, ph.energy, %d));",
                o.type == MXB_OP_FILTER ? "ph.hit" : "active", p.c_str(), fl);
            break;
        }
        case MXB_OP_GRATING: {
            if (o.pg < 0) return fail("GRATING without a selector block");
            const int kind = baked(o.pg);
            std::string select;
            if (kind == MXB_SEL_ORDERSELECTOR) {
                const int n = baked(o.pg + 1);
                if (n < 1) return fail("OrderSelector without orders");
                if (n <= 32) select = "select_order_fixed<" + std::to_string(n) + ">(" + S(o.pg, 3 + 2 * n) + ", u, psel)";
            }
            if (select.empty()) select = "select_order(" + Bref(o.pg) + ", P.prog, u, energy, bl, psel)";
            out("            {");
            out("            double order = 0, blaze = 0;");
            out("            if (ph.hit) {");
            out("                const double u = %s;", draw(o.s0, 0).c_str());
            out("                bool blocked = false;");
            out("                double trans = 0.0;");
            if (fl & 8) {   // L1 support: block = openfraction, offset of the Si transmission table (global)
                const std::string lb = S(o.c[2], 2);
                out("                blocked = %s > %s[0];", draw(o.s1, 0).c_str(), lb.c_str());
                out("                if (blocked) trans = filter_value(st_sm, PRef<false>{P.prog, (int)%s[1]}, ph.energy, 1);", lb.c_str());
            }
            out("                op_grating(st_sm, ph, %s, %s, %d,", PR(o, 9).c_str(), geom.c_str(), fl);
            const std::string dcol = (fl & 16) ? F(o.c[3]) + "[i]" : std::string("0.0");   // callable d: per-photon column
            out("                           [&](double energy, double bl, double& psel) { return %s; }, order, blaze, blocked, trans, %s);",
                select.c_str(), dcol.c_str());
            out("            }");
            put(o.c[0], "order");
            put(o.c[1], "blaze");
            out("            }");
            break;
        }
        case MXB_OP_DETPIX: {
            out("            {");
            out("            double px = 0, py = 0;");
            out("            if (ph.hit) op_detpix(ph, %s, %d, px, py);", PR(o, (fl & 2) ? 4 : 3).c_str(), fl);
            put(o.c[0], "px");
            put(o.c[1], "py");
            if (o.s0 >= 0 && o.s0 < MXB_MAX_F64_COLS && cols->f64[o.s0]) {
                if (o.pg < 0) return fail("DETPIX image without a header");
                need_hot = true;
                out("            if (ph.hit) accumulate_image(hot, %s, %s, %s, px, py, ph.prob);", F(o.s0).c_str(), S(o.pg, 4).c_str(),
                    row_id(o, (fl & 1) != 0).c_str());
            }
            out("            }");
            break;
        }
        case MXB_OP_ACIS: {
            if (!in_array) return fail("ACIS outside an array");
            const bool img = o.s0 >= 0 && o.s0 < MXB_MAX_F64_COLS && cols->f64[o.s0];
            const std::string gp = S(o.pg, img ? 10 : 6);
            out("            {");
            out("            double o8[8] = {0, 0, 0, 0, 0, 0, 0, 0};");
            out("            if (ph.hit) op_acis(ph, %s, %s, o8);", PR(o, 8).c_str(), gp.c_str());
            for (int k = 0; k < 8; ++k) put(o.c[k], "o8[" + std::to_string(k) + "]");
            if (img) need_hot = true;
            if (img)   // fused detector image (chip pixel convention is 1-based: det_acis.py:33-34)
                out("            if (ph.hit) accumulate_image(hot, %s, %s + 6, %s, o8[0] - 1.0, o8[1] - 1.0, ph.prob);", F(o.s0).c_str(),
                    gp.c_str(), row_id(o, true).c_str());
            out("            }");
            break;
        }
        case MXB_OP_BREWSTER:
            out("            if (ph.hit) op_brewster(st_sm, ph, %s);", PR(o, 21).c_str());
            break;
        case MXB_OP_GENERATE: {
            if (in_array) return fail("GENERATE inside an array");
            const std::string p = S(o.pg, 10);
            out("            {");
            out("            double time = 0, polangle = 0;");
            out("            if (active) op_generate(ph, %s, P.prog, gid, [&](int s) {", p.c_str());
            // the draw source of each slot is known: emit a switch over the (at most four) slots used
            out("                switch (s) {");
            for (int sl : {o.s0, o.s1, o.w14, o.w15})
                if (sl >= 0 && sl < MXB_MAX_SLOTS) out("                case %d: return %s;", sl, draw(sl, 0).c_str());
            out("                default: return 0.0;");
            {
                const int pidx = o.c[1] >= MXB_COL_INIT ? o.c[1] - MXB_COL_INIT : o.c[1];
                const std::string e_in = (fl & 1) ? F(MXB_COL_ENERGY) + "[i]" : std::string("0.0");
                const std::string p_in = ((fl & 2) && has_f(o.c[1])) ? F(pidx) + "[i]" : std::string("0.0");
                out("                } }, %d, %d, %d, %d, time, polangle, %d, %s, %s);", o.s0, o.s1, o.w14, o.w15, fl, e_in.c_str(),
                    p_in.c_str());
            }
            for (int k = 0; k < 4; ++k) {
                const char* val[] = {"time", "polangle", nullptr, nullptr};
                if (!has_f(o.c[k])) continue;
                if ((k == 0 && (fl & 4)) || (k == 1 && (fl & 2))) continue;      // input columns of callable specifications
                const int idx = o.c[k] >= MXB_COL_INIT ? o.c[k] - MXB_COL_INIT : o.c[k];
                if (k < 2) out("            jput(%s, i, active, true, %s);", F(idx).c_str(), val[k]);
                else out("            jput(%s, i, active, true, %s[%d]);", F(idx).c_str(), p.c_str(), 6 + k);
            }
            out("            }");
            break;
        }
        case MXB_OP_POINTING: {
            if (in_array) return fail("POINTING inside an array");
            for (int j = 0; j < 3; ++j)
                if (o.c[j] < 0 || o.c[j] >= MXB_MAX_F64_COLS || !cols->f64[o.c[j]]) return fail("POINTING: missing ra / dec / polangle column");
            const std::string p = S(o.pg, 22);
            out("            if (active) {");
            out("                double ua = 0, zj = 0;");
            if (fl & 1) {
                out("                ua = %s;", draw(o.s0, 0).c_str());
                out("                if (%s[21] > 0.0) zj = %s;", p.c_str(), draw(o.s1, 1).c_str());
            }
            out("                op_pointing(ph, %s, %d, %s[i], %s[i], %s[i], ua, zj);", p.c_str(), fl, F(o.c[0]).c_str(),
                F(o.c[1]).c_str(), F(o.c[2]).c_str());
            out("            }");
            break;
        }
        case MXB_OP_LABCONE:
        case MXB_OP_FARLAB: {
            if (in_array) return fail("source op inside an array");
            if (o.c[0] < 0 || o.c[0] >= MXB_MAX_F64_COLS || !cols->f64[o.c[0]]) return fail("source op: missing polangle column");
            const bool cone = o.type == MXB_OP_LABCONE;
            pos_committed = false;
            out("            if (active) {");
            out("                const double u0 = %s;", draw(o.s0, 0).c_str());
            out("                const double u1 = %s;", draw(o.s1, 0).c_str());
            out("                %s(ph, %s, u0, u1, %s[i]);", cone ? "op_labcone" : "op_farlab", S(o.pg, cone ? 13 : 15).c_str(),
                F(o.c[0]).c_str());
            out("            }");
            break;
        }
        case MXB_OP_CYLINDER: {
            if (in_array) return fail("CYLINDER inside an array");
            geom = S(o.pg, 35);
            new_hit();
            out("            ph.hit = cylinder_intersect(%s, ph.pos, ph.dir, ph.ip, ph.l0, ph.l1) && active;", geom.c_str());
            hit_count(pc);
            break;
        }
        case MXB_OP_QFACTOR:
            out("            if (ph.hit) op_qfactor(st_sm, ph, %s);", PR(o, 3).c_str());
            break;
        case MXB_OP_L2ABS:
            out("            if (ph.hit) op_l2abs(st_sm, ph, %s, %s);", PR(o, 3).c_str(), geom.c_str());
            break;
        case MXB_OP_MLEFF: {
            // tables are searched with per-photon indices: shared memory
            const std::string p = (o.pf >= 0 && in_array) ? Rref(o.pf) : Bref(pr_off(o));
            out("            if (ph.hit) op_mleff(st_sm, ph, %s);", p.c_str());
            break;
        }
        case MXB_OP_APERTURE: {
            const std::string p = PR(o, 17);
            new_hit();
            out("            {");
            out("            bool sel = active;");
            if (o.w14 >= 0) {   // MultiAperture :201-218: injected aperture id, or area-weighted draw
                out("            if (sel) {");
                out("                const double a = %s;", draw(o.w14, 0).c_str());
                if (injected(o.w14)) out("                sel = ((long long)a == %dLL);", o.w15);
                else out("                sel = (a >= %s[15] && a < %s[16]);", p.c_str(), p.c_str());
                out("            }");
            }
            out("            ph.hit = sel;");
            out("            if (sel) {");
            out("                const double u0 = %s;", draw(o.s0, 0).c_str());
            out("                const double u1 = %s;", draw(o.s1, 0).c_str());
            out("                op_aperture(st_sm, ph, %s, %d, u0, u1);", p.c_str(), fl);
            out("            }");
            out("            }");
            hit_count(pc);
            break;
        }
        case MXB_OP_PROPAGATE: {
            const std::string p = PR(o, 1);
            pos_committed = false;
            out("            ph.pos = V3{ph.pos.x + %s[0] * ph.dir.x, ph.pos.y + %s[0] * ph.dir.y, ph.pos.z + %s[0] * ph.dir.z};",
                p.c_str(), p.c_str(), p.c_str());
            break;
        }
        default:
            return fail("op code " + std::to_string(o.type) + " cannot be specialised");
        }
        folded_commit(o);
        return err.empty();
    }

    bool gen_array(int& pc) {
        const Op& o = ops[pc];
        const int bpc = pc;
        const int nF = o.c[0], stride = o.c[1], rows_off = o.c[2], mode = o.c[3], nu = o.c[4], nv = o.c[5];
        const int cs_off = o.c[6], cand_off = o.c[7], n_init = o.s0, init_off = o.s1, end_pc = o.w14;
        if (in_array) return fail("nested arrays");
        if (end_pc <= pc || end_pc >= n_ops || ops[end_pc].type != MXB_OP_ARRAY_END) return fail("ARRAY_BEGIN without ARRAY_END");
        if (nF < 0 || stride < 14 || !range_ok(rows_off, (long long)std::max(nF, 1) * stride)) return fail("array rows outside the blob");
        if (mode == 1 && (!range_ok(cs_off, ((long long)nu * nv + 2) / 2) || !range_ok(cand_off, 0))) return fail("culling grid outside the blob");
        need_blob = true;
        const std::string H = S(o.pg, 18);
        // disjointness certificate (mxb_ops.cuh array_revalidate): mode and table offset are structure
        const int lim_mode = (mode == 1 && range_ok(o.pg, 20)) ? baked(o.pg + 19) : 0;
        const int lim_off = lim_mode == 2 ? baked(o.pg + 18) : 0;
        if (lim_mode == 2 && !range_ok(lim_off, (nF + 2) / 2)) return fail("certificate table outside the blob");
        out("            // ---- ops %d..%d: array of %d facets, stride %d, mode %d (%d x %d cells)", pc, end_pc, nF, stride, mode, nu, nv);
        out("            {");
        out("            ArrayIter it;");
        out("            int row = 0, nhit = 0;");
        out("            bool init_round = true;");
        out("            array_open(it, %s, (B + %d), %d, %d, %d, %d, ph, active, st_sm);", H.c_str(), mode == 1 ? cs_off : 0, nF, mode, nu, nv);
        out("            for (;;) {");
        out("            const bool found = array_search(it, B, %s, (B + %d), (B + %d), %d, %d, %d, %d, %d, ph, row);", H.c_str(),
            mode == 1 ? cs_off : 0, mode == 1 ? cand_off : 0, rows_off, stride, nF, nu, nv);
        out("            ph.hit = found;");
        out("            nhit += found ? 1 : 0;");
        out("            const unsigned m = __ballot_sync(0xffffffffu, found);");
        out("            if (m == 0u) {");
        if (n_init > 0) {
            // the body never ran for this warp: initialise the columns it would have created
            if (!range_ok(init_off, (n_init + 1) / 2)) return fail("array init list outside the blob");
            out("                if (init_round && active) {");
            const int* il = reinterpret_cast<const int*>(W + init_off);
            for (int k = 0; k < n_init; ++k) {
                const int cr = il[k];
                keyi(cr);
                if (cr >= 0) { if (cr < MXB_MAX_F64_COLS && cols->f64[cr]) out("                    %s[i] = nan64();", F(cr).c_str()); }
                else { const int ii = -cr - 2; if (ii >= 0 && ii < MXB_MAX_I64_COLS && cols->i64[ii]) out("                    %s[i] = -1LL;", I(ii).c_str()); }
            }
            out("                }");
        }
        out("                break;");
        out("            }");
        out("            h%d += __popc(m);", bpc);
        in_array = true;
        new_hit();
        geom = "(B + row)";
        for (pc = bpc + 1; pc < end_pc; ++pc) {
            if (ops[pc].type == MXB_OP_ARRAY_BEGIN || ops[pc].type == MXB_OP_ARRAY_END) return fail("nested arrays");
            if (!gen_op(pc)) return false;
        }
        in_array = false;
        new_hit();
        out("            // ---- op %d: ARRAY_END", end_pc);
        {
            // the cone only has to be re-validated when the body can change the photon's direction
            bool redirects = false;
            for (int k = bpc + 1; k < end_pc; ++k) {
                const int t = ops[k].type;
                redirects |= (t == MXB_OP_LENS || t == MXB_OP_RSCATTER || t == MXB_OP_GSCATTER || t == MXB_OP_GRATING ||
                              t == MXB_OP_BREWSTER);
            }
            if (redirects || lim_mode)
                out("            array_revalidate<%s>(it, %s, (B + %d), %d, ph, nhit, row, %d, %d, %d, st_sm);", redirects ? "true" : "false",
                    H.c_str(), lim_off, lim_mode, rows_off, stride, nF);
            else
                out("            // (body does not redirect photons: the candidate list stays valid)");
        }
        out("            init_round = false;               // later rounds only store for photons that hit");
        out("            if (!__any_sync(0xffffffffu, ph.hit && it.cur < it.end)) break;");
        out("            if (!ph.hit) it.cur = it.end;     // lanes that found nothing are done with this array");
        out("            }");
        out("            if (nhit >= 2) count_status(st_sm, MXB_ST_MULTI_HIT);");
        out("            ph.hit = false;");
        out("            }");
        pc = end_pc;
        return err.empty();
    }

    bool run() {
        if (!parse()) return false;
        staged = true;   // decided after the body is known (needs the static shared memory of the hot-pixel cache)
        geom = "SRef{P.s}";
        for (int pc = 0; pc < n_ops; ++pc) {
            const Op& o = ops[pc];
            if (o.type == MXB_OP_ARRAY_BEGIN) { if (!gen_array(pc)) return false; }
            else if (o.type == MXB_OP_ARRAY_END) return fail("ARRAY_END without ARRAY_BEGIN");
            else if (o.type == MXB_OP_END) break;
            else if (!gen_op(pc)) return false;
        }
        if (!err.empty()) return false;
        // one CTA per SM may use 227 KB of dynamic shared memory minus the kernel's static arrays
        staged = (size_t)stage_words * 8 <= (size_t)(232448 - 1024 - (need_hot ? 8192 : 0));
        keyi(need_blob ? 1 : 0);
        if (!emit) return true;
        std::string body;
        body.swap(src);         // body holds the op code; src receives the preamble
        if ((17 + fmap.size() + imap.size() + dmap.size() + smap.size()) * 8 > 32000) return fail("program has too many scalar parameters");
        out("// generated by libmxb (mxb_jit.cpp): specialised driver of one element program");
        {
            // array bodies that redirect photons more than once (two gratings, grating + scatter ...) make steep rays
            // common: every warp then walks the footprint scan with a lane or two, and its code should stay small
            bool compact_scan = false;
            for (int pc = 0; pc < n_ops; ++pc) {
                if (ops[pc].type != MXB_OP_ARRAY_BEGIN) continue;
                int redirecting = 0;
                for (int k = pc + 1; k < n_ops && k < ops[pc].w14; ++k) {
                    const int t = ops[k].type;
                    redirecting += (t == MXB_OP_GRATING || t == MXB_OP_GSCATTER || t == MXB_OP_RSCATTER || t == MXB_OP_LENS ||
                                    t == MXB_OP_BREWSTER) ? 1 : 0;
                }
                compact_scan |= redirecting >= 2;
            }
            if (compact_scan) out("#define MXB_SCAN_UNROLL1 1");
        }
        out("#include \"mxb_ops.cuh\"");
        out("using namespace mxb;");
        out("#ifndef JIT_THREADS\n#define JIT_THREADS 640\n#endif");
        out("#ifndef JIT_MINBLOCKS\n#define JIT_MINBLOCKS 1\n#endif");
        out("#ifndef JIT_PREFETCH\n#define JIT_PREFETCH 2\n#endif");
        // photon index type: launches are sliced below 2^31 photons (mxbjit::launch), so plane addresses are one
        // IMAD.WIDE.U32 off the pointer in the constant bank instead of 64-bit shift/add chains
        out("#ifndef JIT_IDX32\n#define JIT_IDX32 %d\n#endif", idx64 ? 0 : 1);
        out("#if JIT_IDX32\ntypedef unsigned idx_t;\n#else\ntypedef long long idx_t;\n#endif");
        out("#define JIT_STAGE_WORDS %d", need_blob && staged ? stage_words : 0);
        {   // per-warp TMA input pipeline (mxb_ops.cuh InputPipe) when its buffers fit beside the staged program
            const long long stage_b = need_blob && staged ? (long long)stage_words * 8 : 0;
            pipe = pipe_wanted && stage_b + (long long)(threads / 32) * MXB_PIPE_BYTES_PER_WARP <= 225 * 1024;
            out("#define JIT_PIPE %d", pipe ? 1 : 0);
        }
        out("struct JitParams {");
        out("    long long n, id0, flags;");
        out("    unsigned long long seed;");
        out("    unsigned long long* status;");
        out("    const double* prog;");
        out("    const double* in[11];  // core planes the photons are read from (== f[0..10] in place)");
        out("    double* f[%d];", (int)std::max<size_t>(fmap.size(), 1));
        out("    long long* i[%d];", (int)std::max<size_t>(imap.size(), 1));
        out("    const double* d[%d];", (int)std::max<size_t>(dmap.size(), 1));
        out("    double s[%d];", (int)std::max<size_t>(smap.size(), 1));
        out("};");
        out("MXB_DEV void jput(double* col, idx_t i, bool cond, bool hit, double v) { if (cond) st_global(col + i, hit ? v : nan64()); }");
        out("MXB_DEV void jput_id(long long* col, idx_t i, bool cond, bool hit, long long v) { if (cond) st_global(col + i, hit ? v : -1LL); }");
        out("extern \"C\" __global__ void __launch_bounds__(JIT_THREADS, JIT_MINBLOCKS)");
        out("mxb_jit_kernel(const __grid_constant__ JitParams P) {");
        out("    __shared__ unsigned long long st_sm[MXB_ST_OPHITS];");
        out("    const int tid = threadIdx.x;");
        out("#if JIT_STAGE_WORDS > 0");
        out("    __shared__ __align__(8) uint64_t bar;");
        out("    if (tid == 0) mbar_init(&bar, 1);");
        out("    __syncthreads();");
        out("    if (tid == 0) {   // TMA bulk copy of the staged part of the blob (facet rows, grids, tables)");
        out("        const uint32_t bytes = (uint32_t)JIT_STAGE_WORDS * 8u;");
        out("        mbar_expect_tx(&bar, bytes);");
        out("        for (uint32_t off = 0; off < bytes; off += 32768u)");
        out("            bulk_g2s(reinterpret_cast<char*>(g_smem) + off, reinterpret_cast<const char*>(P.prog) + off,");
        out("                     min(bytes - off, 32768u), &bar);");
        out("    }");
        out("#endif");
        out("#if JIT_PIPE");
        out("    __shared__ __align__(8) uint64_t in_bar[JIT_THREADS / 32];");
        out("    if ((tid & 31) == 0) mbar_init(&in_bar[tid >> 5], 1);");
        out("#endif");
        out("    if (tid < MXB_ST_OPHITS) st_sm[tid] = 0ULL;");
        if (need_hot) {
            out("    __shared__ unsigned long long hot_keys[MXB_HOT_SLOTS];");
            out("    __shared__ double hot_vals[MXB_HOT_SLOTS];");
            out("    const HotCache hot{hot_keys, hot_vals};");
            out("    hot_init(hot, tid, JIT_THREADS);");
        }
        out("#if JIT_STAGE_WORDS > 0");
        out("    mbar_wait(&bar, 0);");
        out("#endif");
        out("    __syncthreads();");
        out("    typedef PRef<(JIT_STAGE_WORDS > 0)> Ref;");
        out("    const Ref B{P.prog, 0};");
        out("    (void)B;");
        for (int pc = 0; pc < n_ops; ++pc) {
            const int t = ops[pc].type;
            if (t == MXB_OP_PLANE || t == MXB_OP_LOADHIT || t == MXB_OP_APERTURE || t == MXB_OP_ARRAY_BEGIN || t == MXB_OP_CYLINDER) out("    unsigned h%d = 0u;", pc);
        }
        out("    const double kNaN = nan64();");
        out("    const idx_t stride = (idx_t)gridDim.x * JIT_THREADS;");
        out("    const idx_t n_ph = (idx_t)P.n;");
        out("    const int lane = tid & 31;");
        out("    const bool tma_ok = JIT_PIPE && (P.flags & 1);   // host: all 11 core planes are 16-byte aligned");
        out("    InputPipe pipe;");
        out("#if JIT_PIPE");
        out("    pipe.buf = g_smem + JIT_STAGE_WORDS + (tid >> 5) * MXB_PIPE_WORDS_PER_WARP;");
        out("    pipe.bar = &in_bar[tid >> 5];");
        out("#else");
        out("    pipe.buf = nullptr;");
        out("    pipe.bar = nullptr;");
        out("#endif");
        out("    idx_t base = (idx_t)blockIdx.x * JIT_THREADS + (tid & ~31);   // whole warps");
        out("    pipe_start(pipe, P.in, base, P.n, tma_ok, lane);");
        out("    for (; base < n_ph; base += stride) {");
        out("        const idx_t i = base + lane;");
        out("        const bool active = i < n_ph;");
        out("        const unsigned long long gid = (unsigned long long)(P.id0 + i);");
        out("        (void)gid;");
        out("        Photon ph;");
        if (born) {
            out("        ph.pos = ph.dir = ph.pol = V3{kNaN, kNaN, kNaN};   // born by the program: nothing is read");
            out("        ph.energy = ph.prob = kNaN;");
        } else {
            out("#if JIT_PIPE");
            out("        pipe_load(pipe, P.in, base, base + stride, P.n, tma_ok, lane, active, ph.pos, ph.dir, ph.pol, ph.energy, ph.prob);");
            out("#else");
            out("        {   // lanes past the end of the batch re-read the last photon (their results are never stored)");
            out("            const idx_t il = active ? i : n_ph - 1;");
            out("            ph.pos = V3{P.in[0][il], P.in[1][il], P.in[2][il]};");
            out("            ph.dir = V3{P.in[3][il], P.in[4][il], P.in[5][il]};");
            out("            ph.pol = V3{P.in[6][il], P.in[7][il], P.in[8][il]};");
            out("            ph.energy = P.in[9][il];");
            out("            ph.prob = P.in[10][il];");
            out("        }");
            out("#endif");
        }
        out("        photon_loaded(ph);");
        out("#if JIT_PREFETCH && !%d", born ? 1 : 0);
#if 1
        out("#if JIT_PREFETCH == 2");
        out("        // next group's inputs -> L2 while this one is traced: the 32 photons of a warp are 256 B = two 128-byte");
        out("        // lines of each of the 11 planes; lane l < 22 fetches line (l & 1) of plane (l >> 1): ONE instruction");
        out("        if (base + stride < n_ph && lane < 2 * MXB_IN_PLANES)");
        out("            asm volatile(\"prefetch.global.L2 [%%0];\" ::\"l\"(P.in[lane >> 1] + base + stride + ((lane & 1) << 4)));");
        out("#else");
#endif
        out("        if (i + stride < n_ph) {   // next group's inputs -> L2 while this one is traced");
        out("#pragma unroll");
        out("            for (int k = 0; k < MXB_IN_PLANES; ++k) asm volatile(\"prefetch.global.L2 [%%0];\" ::\"l\"(P.in[k] + i + stride));");
        out("        }");
        out("#endif");
        out("#endif");
        out("        ph.ip = V3{kNaN, kNaN, kNaN};");
        out("        ph.l0 = ph.l1 = kNaN;");
        out("        {");
        src += body;
        out("        }");
        out("        if (active) {");
        out("            P.f[0][i] = ph.pos.x; P.f[1][i] = ph.pos.y; P.f[2][i] = ph.pos.z;");
        out("            P.f[3][i] = ph.dir.x; P.f[4][i] = ph.dir.y; P.f[5][i] = ph.dir.z;");
        out("            P.f[6][i] = ph.pol.x; P.f[7][i] = ph.pol.y; P.f[8][i] = ph.pol.z;");
        out("            P.f[10][i] = ph.prob;");
        out("            if (P.flags & 2) P.f[9][i] = ph.energy;   // out of place: energy travels too");
        out("        }");
        out("    }");
        out("    if ((tid & 31) == 0) {");
        for (int pc = 0; pc < n_ops; ++pc) {
            const int t = ops[pc].type;
            if (t == MXB_OP_PLANE || t == MXB_OP_LOADHIT || t == MXB_OP_APERTURE || t == MXB_OP_ARRAY_BEGIN || t == MXB_OP_CYLINDER)
                out("        if (h%d) atomicAdd(&P.status[MXB_ST_OPHITS + %d], (unsigned long long)h%d);", pc, pc, pc);
        }
        out("    }");
        out("    __syncthreads();");
        if (need_hot) out("    hot_flush(hot, tid, JIT_THREADS);");
        out("    if (tid < MXB_ST_OPHITS && st_sm[tid]) atomicAdd(&P.status[tid], st_sm[tid]);");
        out("}");
        return true;
    }
};

int env_int(const char* name, int dflt);

// launches of 2^31 photons or more need 64-bit photon indices (set by launch() around its init_gen calls)
thread_local bool tl_idx64 = false;

void init_gen(Gen& g, const double* prog_host, size_t words, const MxbColumns* cols, bool emit) {
    g.idx64 = tl_idx64;
    g.threads = env_int("MXB_JIT_THREADS", 0) > 0 ? env_int("MXB_JIT_THREADS", 0) : 640;
    g.pipe_wanted = env_int("MXB_JIT_PIPE", 0) != 0;
    g.W = prog_host;
    g.words = words;
    g.n_ops = (int)prog_host[2];
    g.stage_words = (int)prog_host[4];
    g.cols = cols;
    g.emit = emit;
    // the 11 core planes are always parameters 0..10
    for (int k = 0; k <= MXB_COL_PROB; ++k) g.F(k);
}

// ---------------------------------------------------------------------------
// compile + cache
// ---------------------------------------------------------------------------
struct Kernel {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t fn = nullptr;
    std::vector<int> fmap, imap, dmap, smap;
    int stage_bytes = 0, smem_bytes = 0, threads = 640, blocks_per_sm = 1, regs = 0;
    bool pipe = false;
    bool need_blob = false;
    std::string hash, origin;
    // per-device launch state, guarded by g_mu (one process may drive several GPUs from several threads)
    uint64_t smem_attr_devs = 0;      // bit d: MaxDynamicSharedMemorySize set on device d
    bool occupancy_known = false;
};

std::mutex g_mu;
std::unordered_map<std::string, std::unique_ptr<Kernel>> g_cache;
int g_mode_override = -1;
thread_local std::string g_info;

constexpr int kSpillBytesTolerated = 32;    // per thread, at the default CTA size

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}

// $MXB_CACHE_DIR, else _jit_cache/ next to libmxb.so when that directory can be created (kernels
// compiled on a build host then travel with the package), else ~/.cache/marxs_b200
std::string cache_dir() {
    if (const char* e = getenv("MXB_CACHE_DIR")) return e;
    static std::string dir;
    static std::once_flag once;
    std::call_once(once, [] {
        Dl_info info;
        if (dladdr((const void*)&cache_dir, &info) && info.dli_fname) {
            std::string p = info.dli_fname;
            const size_t slash = p.rfind('/');
            if (slash != std::string::npos) {
                p = p.substr(0, slash) + "/_jit_cache";
                mkdir(p.c_str(), 0755);
                if (access(p.c_str(), W_OK) == 0) { dir = p; return; }
            }
        }
        const char* home = getenv("HOME");
        dir = std::string(home && *home ? home : "/tmp") + "/.cache/marxs_b200";
    });
    return dir;
}

void mkdirs(const std::string& d) {
    std::string p;
    for (size_t k = 0; k <= d.size(); ++k) {
        if (k == d.size() || d[k] == '/') { if (!p.empty()) mkdir(p.c_str(), 0755); }
        if (k < d.size()) p += d[k];
    }
}

bool read_file(const std::string& path, std::string& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(sz > 0 ? sz : 0);
    const size_t got = sz > 0 ? fread(&out[0], 1, sz, f) : 0;
    fclose(f);
    return got == (size_t)(sz > 0 ? sz : 0) && sz > 0;
}

void write_file(const std::string& path, const std::string& data) {
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    fwrite(data.data(), 1, data.size(), f);
    fclose(f);
    rename(tmp.c_str(), path.c_str());
}

const char* kStdintStub =
    "#pragma once\n"
    "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
    "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;\n";
const char* kStddefStub = "#pragma once\ntypedef unsigned long size_t;\n";

// "ptxas info    : Used N registers, ..." of the kernel in the NVRTC log (-1: not found)
int regs_in_log(const std::string& log) {
    const size_t p = log.find(" registers");
    if (p == std::string::npos) return -1;
    size_t b = p;
    while (b > 0 && isdigit((unsigned char)log[b - 1])) --b;
    return b < p ? atoi(log.substr(b, p - b).c_str()) : -1;
}

// ptxas -v line of the kernel in the NVRTC log: "... N bytes spill stores, M bytes spill loads"
int spill_bytes_in_log(const std::string& log) {
    const size_t p = log.find(" bytes spill stores");
    if (p == std::string::npos) return -1;
    size_t b = p;
    while (b > 0 && isdigit((unsigned char)log[b - 1])) --b;
    return atoi(log.substr(b, p - b).c_str());
}

int compile(const std::string& source, bool fast_build, const std::string& hash, int threads, std::string& cubin,
            int* spill_bytes, std::string* err, int* regs_used = nullptr) {
    Nvrtc& N = nvrtc();
    if (!N.h) { *err = "NVRTC unavailable: " + N.why; return MXB_EJIT; }
    const std::string dir = cache_dir();
    const std::string cu_path = dir + "/mxb_jit_" + hash + ".cu";
    const char* headers[] = {kSrcMxbH, kSrcDeviceCuh, kSrcOpsCuh, kStdintStub, kStddefStub};
    const char* names[] = {"mxb.h", "mxb_device.cuh", "mxb_ops.cuh", "stdint.h", "stddef.h"};
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device"};
    // with the disk cache the embedded headers are written next to the kernels and included from
    // there, so -lineinfo points at real files (ncu --import-source, cuobjdump)
    int n_mem_headers = 5;
    if (env_int("MXB_JIT_DISK_CACHE", 1) != 0) {
        uint64_t hh = fnv1a(kSrcMxbH, strlen(kSrcMxbH));
        hh = fnv1a(kSrcDeviceCuh, strlen(kSrcDeviceCuh), hh);
        hh = fnv1a(kSrcOpsCuh, strlen(kSrcOpsCuh), hh);
        char hs[32];
        snprintf(hs, sizeof(hs), "%016llx", (unsigned long long)hh);
        const std::string inc = dir + "/include_" + hs;
        mkdirs(inc);
        std::string probe;
        if (!read_file(inc + "/mxb_ops.cuh", probe)) {
            for (int k = 0; k < 5; ++k) write_file(inc + "/" + names[k], headers[k]);
        }
        if (read_file(inc + "/mxb_ops.cuh", probe) && probe == kSrcOpsCuh) {
            opts.push_back("-I" + inc);
            n_mem_headers = 0;
        }
    }
    nvrtcProgram prog = nullptr;
    int rc = N.CreateProgram(&prog, source.c_str(), cu_path.c_str(), n_mem_headers, n_mem_headers ? headers : nullptr,
                             n_mem_headers ? names : nullptr);
    if (rc) { *err = std::string("nvrtcCreateProgram: ") + N.GetErrorString(rc); return MXB_EJIT; }
    if (fast_build) opts.push_back("-DMXB_FAST"); else opts.push_back("--fmad=false");
    opts.push_back("-DJIT_THREADS=" + std::to_string(threads));
    opts.push_back("--ptxas-options=-v");
    opts.push_back("-DJIT_MINBLOCKS=" + std::to_string(env_int("MXB_JIT_MINBLOCKS", 1)));
    opts.push_back("-DJIT_PREFETCH=" + std::to_string(env_int("MXB_JIT_PREFETCH", 2)));
    if (const char* extra = getenv("MXB_JIT_DEFINES")) {      // experiments: space separated -D options
        std::string e(extra);
        size_t pos = 0;
        while (pos < e.size()) {
            size_t sp = e.find(' ', pos);
            if (sp == std::string::npos) sp = e.size();
            if (sp > pos) opts.push_back(e.substr(pos, sp - pos));
            pos = sp + 1;
        }
    }
    if (env_int("MXB_JIT_MAXREG", 0) > 0) opts.push_back("--maxrregcount=" + std::to_string(env_int("MXB_JIT_MAXREG", 0)));
    std::vector<const char*> copts;
    for (auto& o : opts) copts.push_back(o.c_str());
    rc = N.CompileProgram(prog, (int)copts.size(), copts.data());
    if (rc) {
        size_t ls = 0;
        N.GetProgramLogSize(prog, &ls);
        std::string log(ls, '\0');
        if (ls) N.GetProgramLog(prog, &log[0]);
        *err = std::string("NVRTC compile failed: ") + N.GetErrorString(rc) + "\n" + log;
        N.DestroyProgram(&prog);
        return MXB_EJIT;
    }
    if (spill_bytes) {
        size_t ls = 0;
        N.GetProgramLogSize(prog, &ls);
        std::string log(ls, '\0');
        if (ls) N.GetProgramLog(prog, &log[0]);
        *spill_bytes = spill_bytes_in_log(log);
        if (regs_used) *regs_used = regs_in_log(log);
    }
    size_t cs = 0;
    N.GetCUBINSize(prog, &cs);
    cubin.resize(cs);
    if (cs) N.GetCUBIN(prog, &cubin[0]);
    N.DestroyProgram(&prog);
    if (!cs) { *err = "NVRTC produced no cubin"; return MXB_EJIT; }
    return MXB_OK;
}

std::string options_tag(bool fast_build) {
    char b[160];
    snprintf(b, sizeof(b), "%s t%d b%d r%d p%d f%d g%d", fast_build ? "fast" : "strict", env_int("MXB_JIT_THREADS", 0),
             env_int("MXB_JIT_MINBLOCKS", 1), env_int("MXB_JIT_MAXREG", 0), env_int("MXB_JIT_PIPE", 0),
             env_int("MXB_JIT_PREFETCH", 2), env_int("MXB_JIT_GROW", 1));
    std::string tag(b);
    if (const char* extra = getenv("MXB_JIT_DEFINES")) tag += std::string(" ") + extra;
    return tag;
}

// source -> cubin through the disk cache
int get_cubin(Gen& g, bool fast_build, std::string& cubin, std::string& hash, std::string& origin, std::string* err) {
    Nvrtc& N = nvrtc();
    if (!N.h) { *err = "NVRTC unavailable: " + N.why; return MXB_EJIT; }
    // disk cache key: generated source + embedded headers + options + compiler version
    uint64_t h = fnv1a(g.src.data(), g.src.size());
    h = fnv1a(kSrcMxbH, strlen(kSrcMxbH), h);
    h = fnv1a(kSrcDeviceCuh, strlen(kSrcDeviceCuh), h);
    h = fnv1a(kSrcOpsCuh, strlen(kSrcOpsCuh), h);
    const std::string tag = options_tag(fast_build) + " nvrtc" + std::to_string(N.major) + "." + std::to_string(N.minor);
    h = fnv1a(tag.data(), tag.size(), h);
    char hs[32];
    snprintf(hs, sizeof(hs), "%016llx", (unsigned long long)h);
    const std::string dir = cache_dir();
    const bool use_disk = env_int("MXB_JIT_DISK_CACHE", 1) != 0;
    origin = "compiled";
    hash = hs;
    const std::string cubin_path = dir + "/mxb_jit_" + hs + ".cubin";
    if (use_disk && read_file(cubin_path, cubin)) origin = "disk cache";
    else {
        cubin.clear();
        // CTA size: $MXB_JIT_THREADS, else 640 (20 warps at <= 96 registers) unless ptxas reports spills at that
        // register budget - long element stacks (C3: five-layer CAT facets) then run faster with 512 threads
        // and 128 registers (measured 14 % on B200).  The launch reads the size back from the cubin
        // (maxThreadsPerBlock = the __launch_bounds__ the kernel was compiled with).
        const int fixed = env_int("MXB_JIT_THREADS", 0);
        int spill = -1, regs = -1;
        int rc = compile(g.src, fast_build, hs, fixed > 0 ? fixed : 640, cubin, &spill, err, &regs);
        if (rc) return rc;
        if (fixed <= 0 && spill > kSpillBytesTolerated) {
            std::string alt, alt_err;
            int alt_spill = -1;
            if (compile(g.src, fast_build, hs, 512, alt, &alt_spill, &alt_err) == MXB_OK && alt_spill >= 0 && alt_spill < spill)
                cubin.swap(alt);
        } else if (fixed <= 0 && spill == 0 && regs > 0 && regs <= 80 && env_int("MXB_JIT_GROW", 1)) {
            // short programs (C1, C4: <= 80 registers at 640 threads) leave register file unused: each SM
            // sub-partition has 16384 registers, so 80 registers allow 6 warps (768 threads) and 64 allow 8
            // (1024).  Recompile with the larger __launch_bounds__ and keep it if it still does not spill.
            const int want = regs <= 64 ? 1024 : 768;
            std::string alt, alt_err;
            int alt_spill = -1;
            if (compile(g.src, fast_build, hs, want, alt, &alt_spill, &alt_err) == MXB_OK && alt_spill == 0) cubin.swap(alt);
        }
        if (use_disk) {
            mkdirs(dir);
            write_file(dir + "/mxb_jit_" + hs + ".cu", g.src);
            write_file(cubin_path, cubin);
        }
    }
    return MXB_OK;
}

int get_kernel(Gen& keygen, const double* prog_host, size_t words, const MxbColumns* cols, bool fast_build,
               Kernel** out, std::string* err) {
    std::string key = keygen.key;
    key += options_tag(fast_build);
    if (tl_idx64) key += " i64";
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) { *out = it->second.get(); return MXB_OK; }
    Gen g;
    init_gen(g, prog_host, words, cols, true);
    if (!g.run()) { *err = "specialisation failed: " + g.err; return MXB_EJIT; }
    std::string cubin, hs, origin;
    const int rc0 = get_cubin(g, fast_build, cubin, hs, origin, err);
    if (rc0) return rc0;
    const std::string cubin_path = cache_dir() + "/mxb_jit_" + hs + ".cubin";
    std::unique_ptr<Kernel> k(new Kernel());
    cudaError_t ce = cudaLibraryLoadData(&k->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        if (origin == "disk cache") unlink(cubin_path.c_str());
        *err = std::string("cudaLibraryLoadData: ") + cudaGetErrorString(ce);
        return MXB_ECUDA;
    }
    ce = cudaLibraryGetKernel(&k->fn, k->lib, "mxb_jit_kernel");
    if (ce != cudaSuccess) { cudaGetLastError(); *err = std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(ce); return MXB_ECUDA; }
    k->fmap = g.fmap; k->imap = g.imap; k->dmap = g.dmap; k->smap = g.smap;
    k->need_blob = g.need_blob;
    k->stage_bytes = (g.need_blob && g.staged) ? g.stage_words * 8 : 0;
    k->pipe = g.pipe;
    k->smem_bytes = k->stage_bytes + (g.pipe ? (g.threads / 32) * MXB_PIPE_BYTES_PER_WARP : 0);
    k->threads = env_int("MXB_JIT_THREADS", 0) > 0 ? env_int("MXB_JIT_THREADS", 0) : 640;
    k->hash = hs;
    k->origin = origin;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, (const void*)k->fn) == cudaSuccess) {
        k->regs = fa.numRegs;
        if (fa.maxThreadsPerBlock > 0 && fa.maxThreadsPerBlock <= 1024) k->threads = fa.maxThreadsPerBlock;   // = JIT_THREADS
    } else cudaGetLastError();
    *out = k.get();
    g_cache[key] = std::move(k);
    return MXB_OK;
}

int sm_count_of(int dev) {
    static int sms[64] = {0};
    if (dev < 0 || dev >= 64) return 148;
    if (!sms[dev]) {
        if (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); sms[dev] = 148; }
    }
    return sms[dev];
}

}  // namespace

Mode mode() {
    if (g_mode_override >= 0) return (Mode)g_mode_override;
    const char* e = getenv("MXB_JIT");
    if (!e || !*e || !strcmp(e, "auto")) return kAuto;
    if (!strcmp(e, "0") || !strcmp(e, "off")) return kOff;
    return kForce;
}
void set_mode(int m) { g_mode_override = (m < 0 || m > 2) ? -1 : m; }
long long auto_threshold() { return env_int("MXB_JIT_MIN_PHOTONS", 1 << 17); }
const std::string& last_info() { return g_info; }

long long compile_only(const double* prog_host, size_t words, const MxbColumns* cols, bool fast_build,
                       std::string* info, std::string* err) {
    Gen g;
    init_gen(g, prog_host, words, cols, true);
    if (!g.run()) { *err = "specialisation failed: " + g.err; return MXB_EJIT; }
    std::string cubin, hs, origin;
    const int rc = get_cubin(g, fast_build, cubin, hs, origin, err);
    if (rc) return rc;
    *info = "jit " + hs + " (" + origin + ")";
    return (long long)cubin.size();
}

std::string source_for(const double* prog_host, size_t words, const MxbColumns* cols, std::string* err) {
    Gen g;
    init_gen(g, prog_host, words, cols, true);
    if (!g.run()) { *err = g.err; return std::string(); }
    return g.src;
}

int launch(const double* prog_dev, const double* prog_host, size_t words, int n_ops, int stage_words,
           const double* const* src, const MxbColumns* cols, int64_t n, int64_t id0, uint64_t seed,
           unsigned long long* status, cudaStream_t stream, bool fast_build, std::string* err, bool* unavailable) {
    (void)n_ops; (void)stage_words;
    *unavailable = false;
    if (!nvrtc().h) { *unavailable = true; *err = "NVRTC unavailable: " + nvrtc().why; return MXB_EJIT; }
    struct Idx64Scope {
        explicit Idx64Scope(bool v) { tl_idx64 = v; }
        ~Idx64Scope() { tl_idx64 = false; }
    } idx_scope(n >= (1LL << 31));
    Gen kg;
    init_gen(kg, prog_host, words, cols, false);
    if (!kg.run()) { *err = "specialisation failed: " + kg.err; return MXB_EJIT; }
    Kernel* k = nullptr;
    int rc = get_kernel(kg, prog_host, words, cols, fast_build, &k, err);
    if (rc) return rc;
    int dev = 0;
    cudaGetDevice(&dev);
    int blocks_per_sm = 1;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        const uint64_t dev_bit = (dev >= 0 && dev < 64) ? (1ULL << dev) : 0ULL;
        if (k->smem_bytes > 48 * 1024 && !(k->smem_attr_devs & dev_bit)) {
            cudaError_t ce = cudaFuncSetAttribute((const void*)k->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k->smem_bytes);
            if (ce != cudaSuccess) { cudaGetLastError(); *err = std::string("cudaFuncSetAttribute(smem): ") + cudaGetErrorString(ce); return MXB_ECUDA; }
            k->smem_attr_devs |= dev_bit;
        }
        if (!k->occupancy_known) {
            k->occupancy_known = true;
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*)k->fn, k->threads, k->smem_bytes) == cudaSuccess && nb > 0)
                k->blocks_per_sm = nb;
            else { cudaGetLastError(); k->blocks_per_sm = 1; }
        }
        blocks_per_sm = k->blocks_per_sm;
    }
    // parameter block: n id0 seed status prog f[] i[] d[] s[]   (all 8-byte words)
    std::vector<uint64_t> pw;
    pw.reserve(17 + k->fmap.size() + k->imap.size() + k->dmap.size() + k->smap.size() + 4);
    pw.push_back((uint64_t)n);
    pw.push_back((uint64_t)id0);
    uint64_t flags = 1;   // bit 0: the 11 source planes are 16-byte aligned (TMA input pipeline)
    for (int c = 0; c <= MXB_COL_PROB; ++c)
        if ((uintptr_t)(src ? src[c] : cols->f64[c]) & 15) flags = 0;
    if (src && src[MXB_COL_ENERGY] != cols->f64[MXB_COL_ENERGY]) flags |= 2;   // bit 1: store energy
    if ((int)prog_host[5] & 1) flags = (flags | 2) & ~1ULL;                     // born photons: energy is a result, no input pipe
    pw.push_back(flags);
    pw.push_back(seed);
    pw.push_back((uint64_t)(uintptr_t)status);
    pw.push_back((uint64_t)(uintptr_t)prog_dev);
    for (int c = 0; c <= MXB_COL_PROB; ++c) pw.push_back((uint64_t)(uintptr_t)(src ? src[c] : cols->f64[c]));
    for (int c : k->fmap) pw.push_back((uint64_t)(uintptr_t)cols->f64[c]);
    if (k->fmap.empty()) pw.push_back(0);
    for (int c : k->imap) pw.push_back((uint64_t)(uintptr_t)cols->i64[c]);
    if (k->imap.empty()) pw.push_back(0);
    for (int c : k->dmap) pw.push_back((uint64_t)(uintptr_t)cols->draws[c]);
    if (k->dmap.empty()) pw.push_back(0);
    for (int o : k->smap) { uint64_t u; memcpy(&u, prog_host + o, 8); pw.push_back(u); }
    if (k->smap.empty()) pw.push_back(0);
    long long blocks = (n + k->threads - 1) / k->threads;
    const long long cap = (long long)sm_count_of(dev) * blocks_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    void* args[] = {pw.data()};
    cudaError_t ce = cudaLaunchKernel((const void*)k->fn, dim3((unsigned)blocks), dim3((unsigned)k->threads), args,
                                      (size_t)k->smem_bytes, stream);
    if (ce != cudaSuccess) { cudaGetLastError(); *err = std::string("cudaLaunchKernel(jit): ") + cudaGetErrorString(ce); return MXB_ECUDA; }
    char b[256];
    snprintf(b, sizeof(b), "jit %s (%s): %d regs, %d threads x %lld CTAs (%d/SM), %d B staged, %d scalars, input pipe %s",
             k->hash.c_str(), k->origin.c_str(), k->regs, k->threads, blocks, blocks_per_sm, k->stage_bytes,
             (int)k->smap.size(), k->pipe ? ((flags & 1) ? "on" : "off (unaligned planes)") : "off");
    g_info = b;
    return MXB_OK;
}

}  // namespace mxbjit
