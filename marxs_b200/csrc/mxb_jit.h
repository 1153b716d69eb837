// mxb_jit.h — per-program kernel specialisation (host side, internal to libmxb).
//
// The op list of a program blob is turned into a straight-line CUDA driver over the op bodies of
// mxb_ops.cuh, compiled once with NVRTC for sm_100a and cached (in memory and on disk) under a key
// of the program's STRUCTURE: op codes, flags, offsets, which columns exist, which draws are
// injected.  Numeric parameters (geometry, tables) are NOT baked, so re-positioning elements
// (tolerancing loops) never recompiles.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

#include "../../include/mxb.h"

namespace mxbjit {

enum Mode { kOff = 0, kAuto = 1, kForce = 2 };

// mode from MXB_JIT (0 | auto | 1) unless overridden with set_mode
Mode mode();
void set_mode(int m);          // -1: back to the environment
long long auto_threshold();    // photons per launch from which `auto` specialises

// Launch the specialised kernel of this program.  Returns MXB_OK, or an MXB_E* code with *err set.
// `unavailable` is set when NVRTC cannot be loaded at all (auto mode then uses the interpreter).
int launch(const double* prog_dev, const double* prog_host, size_t words, int n_ops, int stage_words,
           const double* const* src, const MxbColumns* cols, int64_t n, int64_t id0, uint64_t seed,
           unsigned long long* status, cudaStream_t stream, bool fast_build, std::string* err, bool* unavailable);

// CUDA source of the specialised kernel (for inspection / profiles); empty + *err on failure
std::string source_for(const double* prog_host, size_t words, const MxbColumns* cols, std::string* err);

// compile (or fetch from the disk cache) without loading: needs NVRTC but no GPU.  Returns the cubin size.
long long compile_only(const double* prog_host, size_t words, const MxbColumns* cols, bool fast_build,
                       std::string* info, std::string* err);

// one-line description of the most recent specialised launch on this thread
const std::string& last_info();

}  // namespace mxbjit
