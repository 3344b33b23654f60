// cps_lm_api.cu -- host side of the batched LM fits: context, launch geometry, the C ABI of
// include/cps_lm.h.  Compiled with -fmad=false together with the kernels (see cps_lm_kernel.cuh).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda.h>  // types of the one driver entry point the asynchronous device entry uses (resolved at run time)

#include "../../include/cps_lm.h"
#include "cps_common.h"
#include "cps_lm_kernel.cuh"
#include <cstdio>
#include "cps_lm_staged.cuh"
#include "cps_lm_dif_staged.cuh"
#include "cps_lm_image8.cuh"

namespace cps {
static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
}  // namespace cps

// One pipeline slot: a stream with its own fit queue, secant-Jacobian workspace and staging block, so
// that the host-buffer entry points can copy chunk i+1 while chunk i computes.
struct cps_lm_slot {
  cudaStream_t stream = nullptr;
  unsigned int* d_queue = nullptr;
  double* d_jstore = nullptr;
  size_t jstore_bytes = 0;
  char* d_stage = nullptr;
  size_t stage_bytes = 0;
  std::vector<long long> off0;  // rebased offsets of the chunk in flight (must outlive the async copy)
  // staged (phase-split) der path: per-fit state, compaction lists, counters
  char* d_staged = nullptr;
  size_t staged_bytes = 0;
  unsigned int* h_count = nullptr;  // pinned: the number of fits that go into the next round
  cudaStream_t side = nullptr;      // staged dif: the finite-difference builds run beside the update builds
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

constexpr int kSlots = 2;
constexpr int kChunkFits = 32768;  // fits per pipelined chunk of the host-buffer entry points
constexpr int kStagedTailFits = 16384;  // staged der: the last fits (at most this many and B/2) go to the warp-per-fit kernel
constexpr int kImage8ThreadMinFits = 4096;  // m = 8: a thread per fit from this many fits per call (measured crossover)
constexpr int kStagedMinFits = 8192;   // der, m = 19: staged shape from this many fits per call (measured crossover ~6 k)
constexpr int kDifStagedMinFits = 14336;  // dif, m = 19 (forward differences): staged shape from this many fits per call (measured crossover ~12 k)
constexpr int kDifChunkFits = 131072;    // host-buffer entry, staged dif: fits per pipelined chunk (amortises the late rounds)
constexpr size_t kDifMaxJBytes = (size_t)48 << 30;  // stored secant Jacobians of one staged dif pass (larger batches: several passes)

struct cps_lm_ctx {
  int device = 0;
  int sm_count = 0;
  int max_smem_optin = 0;
  cps_lm_slot slot[kSlots];
  int last_grid = 0, last_block = 0, last_smem = 0;
  long long launches = 0;
  int path = 0;          // CPS_LM_PATH_AUTO / _MONOLITHIC / _STAGED
  int arith = 0;         // CPS_ARITH_EXACT / CPS_ARITH_FMA (staged der only)
  int last_rounds = 0;   // rounds of the last staged call
  // optional per-kernel timing of the staged rounds (CUDA events on the launching stream)
  int profiling = 0;
  double prof_ms[4] = {0, 0, 0, 0};     // solve, eval (incl. the initial pass), jac, tail (warp-per-fit take-over)
  long long prof_launches[4] = {0, 0, 0, 0};
  cudaEvent_t ev_last = nullptr;  // end of the last device-entry call (calls on different streams serialise on it)
  long long solve_handovers = 0;  // systems of the last staged der call that left the block-arrow solve for the dense one
  // the staged der call about to be made holds fits of different lengths (frames entry; host entry with ragged
  // offsets): order them by length before they enter the lists (cps_lm_staged.cuh, staged_bucket_*)
  int bucket_hint = 0;
  long long prof_jac_evals = 0, prof_tail_fits = 0;  // Jacobian evaluations done by staged_jac_kernel; fits taken over
  std::vector<cudaEvent_t> ev_pool;
  // asynchronous device entry (cps_lm_set_async): a worker thread runs the staged shape's host loop on a private stream;
  // the caller's stream waits on a sequence number in mapped host memory
  struct AsyncJob {
    int model, variant, B, n_max, itmax;
    double* d_p; const double* d_meas; const double* d_t; const long long* d_off; const int* d_counts;
    bool has_opts; double opts[5];
    double* d_info; int* d_ret; int* d_extra;
    cudaEvent_t ev_in;
    unsigned int seq;
  };
  int async_on = 0;
  std::thread worker;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::deque<AsyncJob> jobs;
  bool stop = false;
  unsigned int seq_issued = 0, seq_done = 0;
  volatile unsigned int* flag_host = nullptr;
  CUdeviceptr flag_dev = 0;
  cudaStream_t async_stream = nullptr;
  int async_rc = 0;
  std::string async_error;
  void* wait_value32 = nullptr;  // cuStreamWaitValue32
};

extern "C" const char* cps_last_error(void) { return cps::g_last_error.c_str(); }

// cps_lm_fast.cu: the staged der kernels once more, with fused multiply-adds (cps_lm_set_arith)
extern "C" int cps_fast_prepare(int jac_smem, int eval_smem);
extern "C" void cps_fast_launch_jac(int two_threads, int grid, int block, int smem, cudaStream_t st, const void* args, int nxt);
extern "C" void cps_fast_launch_eval(int init, int grid, int block, int smem, cudaStream_t st, const void* args, int nxt);

extern "C" void cps_lm_destroy(cps_lm_ctx* c);
extern "C" int cps_lm_create(cps_lm_ctx** out, int device) {
  if (!out) return CPS_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
    cps::set_last_error("no CUDA device: libcps has no CPU path");
    return CPS_ERR_CUDA;
  }
  if (device < 0 || device >= count) return CPS_ERR_INVALID;
  CPS_CUDA(cudaSetDevice(device));
  cps_lm_ctx* c = new cps_lm_ctx();
  c->device = device;
  auto init = [&]() -> int {
    CPS_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CPS_CUDA(cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    for (int i = 0; i < kSlots; ++i) {
      CPS_CUDA(cudaStreamCreateWithFlags(&c->slot[i].stream, cudaStreamNonBlocking));
      CPS_CUDA(cudaMalloc(&c->slot[i].d_queue, sizeof(unsigned int)));
    }
    return CPS_OK;
  };
  const int rc = init();
  if (rc != CPS_OK) {
    cps_lm_destroy(c);
    return rc;
  }
  *out = c;
  return CPS_OK;
}

extern "C" void cps_lm_destroy(cps_lm_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->worker.joinable()) {
    {
      std::lock_guard<std::mutex> lk(c->mu);
      c->stop = true;
    }
    c->cv_job.notify_all();
    c->worker.join();
  }
  if (c->async_stream) cudaStreamDestroy(c->async_stream);
  if (c->flag_host) cudaFreeHost((void*)c->flag_host);
  for (int i = 0; i < kSlots; ++i) {
    cps_lm_slot& s = c->slot[i];
    if (s.stream) cudaStreamSynchronize(s.stream);
    cudaFree(s.d_queue);
    cudaFree(s.d_jstore);
    cudaFree(s.d_stage);
    cudaFree(s.d_staged);
    if (s.h_count) cudaFreeHost(s.h_count);
    if (s.ev_fork) cudaEventDestroy(s.ev_fork);
    if (s.ev_join) cudaEventDestroy(s.ev_join);
    if (s.side) cudaStreamDestroy(s.side);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  if (c->ev_last) cudaEventDestroy(c->ev_last);
  delete c;
}

extern "C" int cps_lm_sync(cps_lm_ctx* c) {
  if (!c) return CPS_ERR_INVALID;
  int rc = CPS_OK;
  if (c->async_on) {  // every call handed to the worker has run; the first error among them is reported here
    std::unique_lock<std::mutex> lk(c->mu);
    c->cv_done.wait(lk, [&] { return c->seq_done == c->seq_issued; });
    rc = c->async_rc;
    if (rc != CPS_OK) cps::set_last_error(c->async_error);
    c->async_rc = CPS_OK;
  }
  for (int i = 0; i < kSlots; ++i) CPS_CUDA(cudaStreamSynchronize(c->slot[i].stream));
  return rc;
}

extern "C" int cps_lm_set_arith(cps_lm_ctx* c, int arith) {
  if (!c || (arith != CPS_ARITH_EXACT && arith != CPS_ARITH_FMA)) return CPS_ERR_INVALID;
  c->arith = arith;
  return CPS_OK;
}

extern "C" int cps_lm_set_path(cps_lm_ctx* c, int path) {
  if (!c || path < CPS_LM_PATH_AUTO || path > CPS_LM_PATH_STAGED) return CPS_ERR_INVALID;
  c->path = path;
  return CPS_OK;
}

extern "C" int cps_lm_set_profiling(cps_lm_ctx* c, int on) {
  if (!c) return CPS_ERR_INVALID;
  c->profiling = on ? 1 : 0;
  return CPS_OK;
}

extern "C" int cps_lm_get_profile(cps_lm_ctx* c, double* ms4, long long* launches4, int reset) {
  if (!c) return CPS_ERR_INVALID;
  for (int i = 0; i < 4; ++i) {
    if (ms4) ms4[i] = c->prof_ms[i];
    if (launches4) launches4[i] = c->prof_launches[i];
    if (reset) { c->prof_ms[i] = 0.0; c->prof_launches[i] = 0; }
  }
  return CPS_OK;
}

extern "C" int cps_lm_get_solve_handovers(cps_lm_ctx* c, long long* systems) {
  if (!c || !systems) return CPS_ERR_INVALID;
  *systems = c->solve_handovers;
  return CPS_OK;
}

extern "C" int cps_lm_get_profile_work(cps_lm_ctx* c, long long* jac_evals, long long* tail_fits, int reset) {
  if (!c) return CPS_ERR_INVALID;
  if (jac_evals) *jac_evals = c->prof_jac_evals;
  if (tail_fits) *tail_fits = c->prof_tail_fits;
  if (reset) { c->prof_jac_evals = 0; c->prof_tail_fits = 0; }
  return CPS_OK;
}

extern "C" int cps_lm_launch_stats(cps_lm_ctx* c, int* grid, int* block, int* smem_bytes, long long* launches) {
  if (!c) return CPS_ERR_INVALID;
  if (grid) *grid = c->last_grid;
  if (block) *block = c->last_block;
  if (smem_bytes) *smem_bytes = c->last_smem;
  if (launches) *launches = c->launches;
  return CPS_OK;
}

namespace {

template <int MODEL, bool DIF>
int launch_fit(cps_lm_ctx* c, cps_lm_slot& sl, cps::LmLaunch& L, cudaStream_t st) {
  using MD = cps::Model<MODEL>;
  auto kern = cps::lm_fit_kernel<MODEL, DIF>;
  const int per_warp = cps::lm_warp_smem_doubles(MD::M, MD::SCRATCH, L.n_max, DIF) * (int)sizeof(double);
  // pick the warps-per-CTA that keeps the most warps resident per SM
  int best_w = 0, best_resident = 0, best_blocks = 0;
  for (int wpc = cps::kWarpsPerCta; wpc >= 1; --wpc) {
    const int smem = per_warp * wpc;
    if (smem > c->max_smem_optin) continue;
    if (!cps::cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                      "cudaFuncSetAttribute"))
      return CPS_ERR_CUDA;
    int blocks = 0;
    if (!cps::cuda_ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kern, wpc * 32, smem),
                      "cudaOccupancyMaxActiveBlocksPerMultiprocessor"))
      return CPS_ERR_CUDA;
    if (blocks * wpc > best_resident) {
      best_resident = blocks * wpc;
      best_w = wpc;
      best_blocks = blocks;
    }
  }
  if (best_w == 0) {
    cps::set_last_error("a fit with this many measurements does not fit in shared memory");
    return CPS_ERR_INVALID;
  }
  if (const char* cap = std::getenv("CPS_LM_MAX_WARPS_PER_SM")) {  // experiment knob: fewer resident warps
    const int c = std::atoi(cap);
    if (c > 0 && best_blocks * best_w > c) best_blocks = std::max(1, c / best_w);
  }
  const int smem = per_warp * best_w;
  CPS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long need_ctas = ((long long)L.B + best_w - 1) / best_w;
  int grid = (int)std::min<long long>(need_ctas, (long long)c->sm_count * best_blocks);
  if (grid < 1) grid = 1;
  {
    const size_t per_warp = (size_t)cps::lm_ws_doubles(MD::M, L.n_max, DIF);
    const size_t need = (size_t)grid * best_w * per_warp * sizeof(double);
    if (need > sl.jstore_bytes) {
      CPS_CUDA(cudaStreamSynchronize(st));
      CPS_CUDA(cudaStreamSynchronize(sl.stream));
      cudaFree(sl.d_jstore);
      sl.d_jstore = nullptr;
      sl.jstore_bytes = 0;
      CPS_CUDA(cudaMalloc(&sl.d_jstore, need));
      sl.jstore_bytes = need;
    }
    L.jstore = sl.d_jstore;
  }
  L.queue = sl.d_queue;
  CPS_CUDA(cudaMemsetAsync(sl.d_queue, 0, sizeof(unsigned int), st));
  kern<<<grid, best_w * 32, smem, st>>>(L);
  CPS_CUDA(cudaGetLastError());
  c->last_grid = grid;
  c->last_block = best_w * 32;
  c->last_smem = smem;
  c->launches += 1;
  return CPS_OK;
}

// dlevmar_der for a large batch of Stereo19 fits as rounds of three phase kernels (cps_lm_staged.cuh).  The
// host only sizes grids: after every round it reads one integer, the number of fits still iterating.
// The batch may arrive in pieces (ranges of the fit index, each guarded by an event recorded on a copy stream):
// a piece joins the rounds as soon as its samples are on the device, so the host-buffer entry points overlap the
// transfer of the batch with the iterations of the fits that are already there.
struct StagedPiece {
  int first, count;
  cudaEvent_t ready;  // nullptr: already resident
};

int launch_staged_der19(cps_lm_ctx* c, cps_lm_slot& sl, cps::LmLaunch& L, cudaStream_t st,
                        const std::vector<StagedPiece>& pieces) {
  using namespace cps;
  const size_t B = (size_t)L.B;
  auto a256 = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const int jac_smem = kJacTableDoubles * kJacThreads * (int)sizeof(double) + (kJacThreads / 32) * (int)sizeof(WarpTiles);
  const int eval_smem = kEvalTableDoubles * kEvalThreads * (int)sizeof(double) + (kEvalThreads / 32) * (int)sizeof(WarpTiles);
  const int solve_smem = kSolveDoubles * kSolveThreads * (int)sizeof(double);
  const int arrow_smem = kArrowSlots * kArrowPitch * (int)sizeof(double);
  // CPS_LM_DENSE_SOLVE=1: the dense work matrix for every system (A/B measurements, tests of the hand-over)
  const bool dense_solve = std::getenv("CPS_LM_DENSE_SOLVE") != nullptr;
  CPS_CUDA(cudaFuncSetAttribute(staged_eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, eval_smem));
  CPS_CUDA(cudaFuncSetAttribute(staged_eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, eval_smem));
  CPS_CUDA(cudaFuncSetAttribute(staged_jac_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, jac_smem));
  CPS_CUDA(cudaFuncSetAttribute(staged_jac_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, jac_smem));
  CPS_CUDA(cudaFuncSetAttribute(staged_solve_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, solve_smem));
  CPS_CUDA(cudaFuncSetAttribute(staged_solve_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, solve_smem));
  CPS_CUDA(cudaFuncSetAttribute(staged_solve_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, arrow_smem));
  int jac_occ = 0, solve_occ = 0, eval_occ = 0, arrow_occ = 0;
  CPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&arrow_occ, staged_solve_kernel<true, false>, kArrowThreads, arrow_smem));
  CPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&jac_occ, staged_jac_kernel<false>, kJacThreads, jac_smem));
  CPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&solve_occ, staged_solve_kernel<false, false>, kSolveThreads, solve_smem));
  CPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&eval_occ, staged_eval_kernel<false>, kEvalThreads, eval_smem));
  const bool fma = c->arith == CPS_ARITH_FMA;
  if (fma && cps_fast_prepare(jac_smem, eval_smem) != 0) {
    cps::set_last_error("the fused-arithmetic kernels do not fit on this device");
    return CPS_ERR_CUDA;
  }
  if (jac_occ < 1 || solve_occ < 1 || eval_occ < 1 || arrow_occ < 1) {
    cps::set_last_error("staged LM kernels do not fit on this device");
    return CPS_ERR_CUDA;
  }
  const int jac_grid_max = c->sm_count * jac_occ;
  const int ws_stride = jac_grid_max * kJacThreads;
  // workspace layout
  size_t o = 0;
  const size_t o_scal = o; o = a256(o + sizeof(StagedScal) * B);
  const size_t o_jtj = o;  o = a256(o + sizeof(double) * kLower19 * B);
  const size_t o_jte = o;  o = a256(o + sizeof(double) * 19 * B);
  const size_t o_pdp = o;  o = a256(o + sizeof(double) * 19 * B);
  const size_t o_tws = o;  o = a256(o + (L.t ? 0 : sizeof(double) * (size_t)(L.n_max / 2) * B));
  const size_t o_ws = o;   o = a256(o + sizeof(double) * kWsSlots * (size_t)ws_stride);
  const size_t o_jl = o;   o = a256(o + sizeof(int) * B);
  const size_t o_el = o;   o = a256(o + sizeof(int) * B);
  const size_t o_s0 = o;   o = a256(o + sizeof(int) * B);
  const size_t o_s1 = o;   o = a256(o + sizeof(int) * B);
  const size_t o_fb = o;   o = a256(o + sizeof(int) * B);
  const size_t o_cnt = o;  o = a256(o + sizeof(unsigned int) * 8);
  // CPS_LM_BUCKETS=0/1 overrides the callers' hint (A/B measurements)
  bool bucket = c->bucket_hint != 0;
  if (const char* e = std::getenv("CPS_LM_BUCKETS")) bucket = std::atoi(e) != 0;
  const int max_chunks = (int)((B + kBucketChunk - 1) / kBucketChunk);
  const size_t o_order = o; o = a256(o + (bucket ? sizeof(int) * B : 0));
  const size_t o_bhist = o; o = a256(o + (bucket ? sizeof(int) * (size_t)kBucketBins * max_chunks : 0));
  if (o > sl.staged_bytes) {
    CPS_CUDA(cudaStreamSynchronize(st));
    CPS_CUDA(cudaStreamSynchronize(sl.stream));
    cudaFree(sl.d_staged);
    sl.d_staged = nullptr;
    sl.staged_bytes = 0;
    CPS_CUDA(cudaMalloc(&sl.d_staged, o));
    sl.staged_bytes = o;
  }
  if (!sl.h_count) CPS_CUDA(cudaMallocHost(&sl.h_count, sizeof(unsigned int) * 4));
  char* d = sl.d_staged;
  StagedArgs A;
  A.L = L;
  A.scal = (StagedScal*)(d + o_scal);
  A.jtj = (double*)(d + o_jtj);
  A.jte = (double*)(d + o_jte);
  A.pdp = (double*)(d + o_pdp);
  A.tws = L.t ? nullptr : (double*)(d + o_tws);
  A.ws = (double*)(d + o_ws);
  A.ws_stride = ws_stride;
  A.jac_list = (int*)(d + o_jl);
  A.eval_list = (int*)(d + o_el);
  A.solve_list[0] = (int*)(d + o_s0);
  A.solve_list[1] = (int*)(d + o_s1);
  A.cnt = (unsigned int*)(d + o_cnt);
  A.fb_list = (int*)(d + o_fb);
  A.force_handover = std::getenv("CPS_LM_FORCE_HANDOVER") ? std::atoi(std::getenv("CPS_LM_FORCE_HANDOVER")) : 0;
  A.init_first = 0;
  A.init_count = 0;
  A.init_order = nullptr;

  auto grid_for = [&](size_t items, int per_block, int cap) {
    size_t gneed = (items + per_block - 1) / per_block;
    if (gneed < 1) gneed = 1;
    return (int)std::min<size_t>(gneed, (size_t)cap);
  };
  CPS_CUDA(cudaMemsetAsync(A.cnt, 0, sizeof(unsigned int) * 8, st));
  // per-kernel timing (cps_lm_set_profiling): an event pair around every launch, read after the round's sync
  struct Span { int kind; cudaEvent_t a, b; };
  std::vector<Span> spans;
  size_t ev_used = 0;
  auto ev_get = [&]() -> cudaEvent_t {
    if (ev_used == c->ev_pool.size()) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
      c->ev_pool.push_back(e);
    }
    return c->ev_pool[ev_used++];
  };
  auto span_begin = [&](int kind) {
    if (!c->profiling) return;
    Span sp{kind, ev_get(), ev_get()};
    if (sp.a && sp.b) { cudaEventRecord(sp.a, st); spans.push_back(sp); }
  };
  auto span_end = [&]() {
    if (!c->profiling || spans.empty()) return;
    cudaEventRecord(spans.back().b, st);
  };
  auto spans_collect = [&]() {
    for (const Span& sp : spans) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
        c->prof_ms[sp.kind] += ms;
        c->prof_launches[sp.kind] += 1;
      }
    }
    spans.clear();
    ev_used = 0;
  };
  size_t next_piece = 0, n_solve = 0;
  int cur = 1, rounds = 0;
  long long jac_evals = 0, tail_taken = 0;
  size_t last_jac = 0;  // jac list length of the last round
  size_t tail_fits = std::min<size_t>(kStagedTailFits, B / 2);
  if (const char* e = std::getenv("CPS_LM_STAGED_TAIL")) tail_fits = (size_t)std::max(0, std::atoi(e));  // 0: never
  if (L.n_max > 40 * kTileRows ||
      lm_warp_smem_doubles(19, Model<19>::SCRATCH, L.n_max, false) * (int)sizeof(double) > c->max_smem_optin)
    tail_fits = 0;  // the warp-per-fit kernel cannot hold fits this long
  for (;;) {
    const int nxt = cur ^ 1;
    if (n_solve > 0) {
      span_begin(0);
      if (dense_solve) {
        staged_solve_kernel<false, false><<<grid_for(n_solve, kSolveThreads, c->sm_count * solve_occ), kSolveThreads, solve_smem, st>>>(A, cur);
        c->launches += 1;
      } else {
        // block-arrow work matrix (96 systems per SM); the systems it hands back (cnt[4]: none, on every batch measured)
        // go through the dense one
        staged_solve_kernel<true, false><<<grid_for(n_solve, kArrowThreads, c->sm_count * arrow_occ), kArrowThreads, arrow_smem, st>>>(A, cur);
        staged_solve_kernel<false, true><<<4, kSolveThreads, solve_smem, st>>>(A, cur);
        c->launches += 2;
      }
      span_end();
    }
    CPS_CUDA(cudaMemsetAsync(A.cnt + 0, 0, sizeof(unsigned int), st));
    CPS_CUDA(cudaMemsetAsync(A.cnt + 2 + nxt, 0, sizeof(unsigned int), st));
    if (n_solve > 0) {
      span_begin(1);
      if (fma) cps_fast_launch_eval(0, grid_for(n_solve, kEvalThreads, c->sm_count * eval_occ), kEvalThreads, eval_smem, st, &A, nxt);
      else staged_eval_kernel<false><<<grid_for(n_solve, kEvalThreads, c->sm_count * eval_occ), kEvalThreads, eval_smem, st>>>(A, nxt);
      span_end();
      c->launches += 1;
    }
    // pieces of the batch that have arrived start here: validation, t, e0, loop-top checks of iteration 0
    size_t added = 0;
    while (next_piece < pieces.size()) {
      const StagedPiece& P = pieces[next_piece];
      if (P.ready) {
        // not there yet: keep iterating if the fits at hand fill the GPU, otherwise let the stream wait for the piece
        // (rounds over a few thousand fits run at the latency floor of the kernels and only add launches)
        if (n_solve + added >= (size_t)jac_grid_max * kJacThreads && cudaEventQuery(P.ready) != cudaSuccess) break;
        CPS_CUDA(cudaStreamWaitEvent(st, P.ready, 0));
      }
      if (P.count > 0) {
        A.init_first = P.first;
        A.init_count = P.count;
        A.init_order = nullptr;
        if (bucket && P.count > 64) {  // the piece's fits ordered by length (three small kernels over the offsets)
          const int nchunks = (P.count + kBucketChunk - 1) / kBucketChunk;
          int* bhist = (int*)(d + o_bhist);
          int* order = (int*)(d + o_order) + P.first;
          const BucketRange R{A.L.off, P.first, P.count};
          staged_bucket_hist_kernel<<<nchunks, 256, 0, st>>>(R, bhist, nchunks);
          staged_bucket_scan_kernel<<<1, 1024, 0, st>>>(bhist, kBucketBins * nchunks);
          staged_bucket_scatter_kernel<<<nchunks, 256, 0, st>>>(R, bhist, nchunks, order);
          A.init_order = order;
          c->launches += 3;
        }
        span_begin(1);
        if (fma) cps_fast_launch_eval(1, grid_for((size_t)P.count, kEvalThreads, c->sm_count * eval_occ), kEvalThreads, eval_smem, st, &A, nxt);
        else staged_eval_kernel<true><<<grid_for((size_t)P.count, kEvalThreads, c->sm_count * eval_occ), kEvalThreads, eval_smem, st>>>(A, nxt);
        span_end();
        c->launches += 1;
        added += (size_t)P.count;
      }
      ++next_piece;
    }
    if (n_solve + added == 0) break;
    // how many fits need a Jacobian: at most all that were solved plus the new arrivals.  In the late rounds most fits
    // are re-solving rejected steps and only a few accepted one; the count falls from round to round, so last round's
    // exact count (published with the round's end) plus a margin picks the two-threads-per-fit variant without a
    // read-back in mid-round.  The kernel takes the real count from the device either way.
    const size_t n_jac = n_solve + added;
    const size_t jac_cap = (size_t)jac_grid_max * kJacThreads;
    const size_t jac_likely = rounds > 0 ? std::min(n_jac, last_jac + last_jac / 4 + added + 64) : n_jac;
    // the grid follows the likely count (blocks without work leave at once, but a snug grid measures a little faster),
    // with a floor that bounds the extra passes should many more fits than expected have accepted a step
    const size_t jac_grid_items = std::max(jac_likely, std::min(n_jac, jac_cap / 4));
    if (n_jac > 0) {
      span_begin(2);
      if (jac_likely * 2 <= jac_cap) {  // few fits: two threads per fit
        if (fma) cps_fast_launch_jac(1, grid_for(jac_grid_items, kJacThreads / 2, jac_grid_max), kJacThreads, jac_smem, st, &A, nxt);
        else staged_jac_kernel<true><<<grid_for(jac_grid_items, kJacThreads / 2, jac_grid_max), kJacThreads, jac_smem, st>>>(A, nxt);
      } else {
        if (fma) cps_fast_launch_jac(0, grid_for(jac_grid_items, kJacThreads, jac_grid_max), kJacThreads, jac_smem, st, &A, nxt);
        else staged_jac_kernel<false><<<grid_for(jac_grid_items, kJacThreads, jac_grid_max), kJacThreads, jac_smem, st>>>(A, nxt);
      }
      span_end();
      c->launches += 1;
    }
    CPS_CUDA(cudaGetLastError());
    staged_round_end_kernel<<<1, 1, 0, st>>>(A.cnt, nxt, sl.h_count);
    c->launches += 1;
    CPS_CUDA(cudaStreamSynchronize(st));
    spans_collect();
    n_solve = sl.h_count[0];
    last_jac = sl.h_count[1];
    jac_evals = sl.h_count[2];
    c->solve_handovers = sl.h_count[3];
    cur = nxt;
    ++rounds;
    if (n_solve == 0 && next_piece == pieces.size()) break;
    if (n_solve > 0 && n_solve <= tail_fits && next_piece == pieces.size()) {
      // the late rounds run at the latency floor of three kernels for a few thousand fits: the persistent
      // warp-per-fit kernel takes the stragglers over where they stand and carries each to its end
      LmLaunch T = L;
      T.B = (int)n_solve;
      T.fit_list = A.solve_list[cur];
      T.resume = A.scal;
      T.resume_jtj = A.jtj;
      T.resume_jte = A.jte;
      tail_taken = (long long)n_solve;
      span_begin(3);
      const int rc = launch_fit<19, false>(c, sl, T, st);
      span_end();
      if (rc != CPS_OK) return rc;
      CPS_CUDA(cudaStreamSynchronize(st));
      spans_collect();
      ++rounds;
      break;
    }
  }
  c->last_grid = jac_grid_max;
  c->last_block = kJacThreads;
  c->last_smem = jac_smem;
  c->last_rounds = rounds;
  if (c->profiling) {
    c->prof_jac_evals += jac_evals;
    c->prof_tail_fits += tail_taken;
  }
  return CPS_OK;
}


// dlevmar_dif for a large batch of Stereo19 fits as rounds of phase kernels (cps_lm_dif_staged.cuh): solve -> eval ->
// build (finite-difference Jacobian | pending secant update, each with the J^T J rebuild).  The host only sizes
// grids: it reads three integers per round.  A batch whose stored Jacobians (n_max x 20 doubles per fit) would not
// fit kDifMaxJBytes is run as several passes over ranges of the fit index.
int launch_staged_dif19_range(cps_lm_ctx* c, cps_lm_slot& sl, cps::LmLaunch L, cudaStream_t st) {
  using namespace cps;
  const size_t B = (size_t)L.B;
  auto a256 = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const int eval_smem = kEvalTableDoubles * kEvalThreads * (int)sizeof(double) + (kEvalThreads / 32) * (int)sizeof(WarpTiles);
  const int solve_smem = kSolveDoubles * kSolveThreads * (int)sizeof(double);
  const int fd_smem = dif_build_smem_doubles<true>() * (int)sizeof(double);
  const int upd_smem = dif_build_smem_doubles<false>() * (int)sizeof(double);
  CPS_CUDA(cudaFuncSetAttribute(dif_eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, eval_smem));
  CPS_CUDA(cudaFuncSetAttribute(dif_eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, eval_smem));
  CPS_CUDA(cudaFuncSetAttribute(dif_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, solve_smem));
  CPS_CUDA(cudaFuncSetAttribute(dif_build_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fd_smem));
  CPS_CUDA(cudaFuncSetAttribute(dif_build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, upd_smem));
  int eval_occ = 0, solve_occ = 0, fd_occ = 0, upd_occ = 0;
  CPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&eval_occ, dif_eval_kernel<false>, kEvalThreads, eval_smem));
  CPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&solve_occ, dif_solve_kernel, kSolveThreads, solve_smem));
  CPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fd_occ, dif_build_kernel<true>, kDifThreads, fd_smem));
  CPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&upd_occ, dif_build_kernel<false>, kDifThreads, upd_smem));
  if (eval_occ < 1 || solve_occ < 1 || fd_occ < 1 || upd_occ < 1) {
    cps::set_last_error("staged dif kernels do not fit on this device");
    return CPS_ERR_CUDA;
  }
  const int build_grid_max = c->sm_count * std::max(fd_occ, upd_occ);
  const int ws_stride = 2 * build_grid_max * kDifFits;  // update kernel's CTAs | finite-difference kernel's CTAs
  const size_t j_stride = (size_t)L.n_max * kDifRow;
  size_t o = 0;
  const size_t o_scal = o; o = a256(o + sizeof(DifScal) * B);
  const size_t o_jtj = o;  o = a256(o + sizeof(double) * kLower19 * B);
  const size_t o_jte = o;  o = a256(o + sizeof(double) * 19 * B);
  const size_t o_pdp = o;  o = a256(o + sizeof(double) * 19 * B);
  const size_t o_dp = o;   o = a256(o + sizeof(double) * 19 * B);
  const size_t o_pold = o; o = a256(o + sizeof(double) * 19 * B);
  const size_t o_tws = o;  o = a256(o + (L.t ? 0 : sizeof(double) * (size_t)(L.n_max / 2) * B));
  const size_t o_ws = o;   o = a256(o + sizeof(double) * kDifSums * (size_t)ws_stride);
  const size_t o_fd = o;   o = a256(o + sizeof(int) * B);
  const size_t o_up = o;   o = a256(o + sizeof(int) * B);
  const size_t o_el = o;   o = a256(o + sizeof(int) * B);
  const size_t o_s0 = o;   o = a256(o + sizeof(int) * B);
  const size_t o_s1 = o;   o = a256(o + sizeof(int) * B);
  const size_t o_cnt = o;  o = a256(o + sizeof(unsigned int) * 4);
  // fits of different lengths enter the lists ordered by length (cps_lm_staged.cuh, staged_bucket_*): a CTA of the build
  // kernels walks as many rows as the longest of its 32 fits has
  bool bucket = c->bucket_hint != 0 && B > 64;
  if (const char* e = std::getenv("CPS_LM_BUCKETS")) bucket = std::atoi(e) != 0 && B > 64;
  const int nchunks = (int)((B + kBucketChunk - 1) / kBucketChunk);
  const size_t o_order = o; o = a256(o + (bucket ? sizeof(int) * B : 0));
  const size_t o_bhist = o; o = a256(o + (bucket ? sizeof(int) * (size_t)kBucketBins * nchunks : 0));
  const size_t o_J = o;    o = a256(o + sizeof(double) * j_stride * B);
  if (o > sl.staged_bytes) {
    CPS_CUDA(cudaStreamSynchronize(st));
    CPS_CUDA(cudaStreamSynchronize(sl.stream));
    cudaFree(sl.d_staged);
    sl.d_staged = nullptr;
    sl.staged_bytes = 0;
    CPS_CUDA(cudaMalloc(&sl.d_staged, o));
    sl.staged_bytes = o;
  }
  if (!sl.h_count) CPS_CUDA(cudaMallocHost(&sl.h_count, sizeof(unsigned int) * 4));
  char* d = sl.d_staged;
  DifArgs A;
  A.L = L;
  A.scal = (DifScal*)(d + o_scal);
  A.jtj = (double*)(d + o_jtj);
  A.jte = (double*)(d + o_jte);
  A.pdp = (double*)(d + o_pdp);
  A.dp = (double*)(d + o_dp);
  A.pold = (double*)(d + o_pold);
  A.tws = L.t ? nullptr : (double*)(d + o_tws);
  A.J = (double*)(d + o_J);
  A.j_stride = j_stride;
  A.ws = (double*)(d + o_ws);
  A.ws_stride = ws_stride;
  A.fd_list = (int*)(d + o_fd);
  A.upd_list = (int*)(d + o_up);
  A.eval_list = (int*)(d + o_el);
  A.solve_list[0] = (int*)(d + o_s0);
  A.solve_list[1] = (int*)(d + o_s1);
  A.cnt = (unsigned int*)(d + o_cnt);
  A.init_first = 0;
  A.init_count = (int)B;
  A.init_order = nullptr;
  auto grid_for = [&](size_t items, int per_block, int cap) {
    size_t gneed = (items + per_block - 1) / per_block;
    if (gneed < 1) gneed = 1;
    return (int)std::min<size_t>(gneed, (size_t)cap);
  };
  struct Span { int kind; cudaEvent_t a, b; };
  std::vector<Span> spans;
  size_t ev_used = 0;
  auto ev_get = [&]() -> cudaEvent_t {
    if (ev_used == c->ev_pool.size()) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
      c->ev_pool.push_back(e);
    }
    return c->ev_pool[ev_used++];
  };
  auto span_begin = [&](int kind) {
    if (!c->profiling) return;
    Span sp{kind, ev_get(), ev_get()};
    if (sp.a && sp.b) { cudaEventRecord(sp.a, st); spans.push_back(sp); }
  };
  auto span_end = [&]() {
    if (!c->profiling || spans.empty()) return;
    cudaEventRecord(spans.back().b, st);
  };
  auto spans_collect = [&]() {
    for (const Span& sp : spans) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
        c->prof_ms[sp.kind] += ms;
        c->prof_launches[sp.kind] += 1;
      }
    }
    spans.clear();
    ev_used = 0;
  };
  if (!sl.side) {
    CPS_CUDA(cudaStreamCreateWithFlags(&sl.side, cudaStreamNonBlocking));
    CPS_CUDA(cudaEventCreateWithFlags(&sl.ev_fork, cudaEventDisableTiming));
    CPS_CUDA(cudaEventCreateWithFlags(&sl.ev_join, cudaEventDisableTiming));
  }
  // The two build kernels work on disjoint fits (and disjoint halves of the work area): the finite-difference one,
  // usually a few thousand fits -- one partial wave -- runs on a side stream beside the update kernel.
  auto builds = [&](size_t n_fd, size_t n_upd, int nxt) {
    const bool fork = n_fd > 0 && n_upd > 0;
    span_begin(2);
    if (n_fd > 0) {
      cudaStream_t s2 = fork ? sl.side : st;
      if (fork) {
        cudaEventRecord(sl.ev_fork, st);
        cudaStreamWaitEvent(sl.side, sl.ev_fork, 0);
      }
      dif_build_kernel<true><<<grid_for(n_fd, kDifFits, c->sm_count * fd_occ), kDifThreads, fd_smem, s2>>>(A, nxt);
      if (fork) cudaEventRecord(sl.ev_join, sl.side);
      c->launches += 1;
    }
    if (n_upd > 0) {
      dif_build_kernel<false><<<grid_for(n_upd, kDifFits, c->sm_count * upd_occ), kDifThreads, upd_smem, st>>>(A, nxt);
      c->launches += 1;
    }
    if (fork) cudaStreamWaitEvent(st, sl.ev_join, 0);
    span_end();
  };
  CPS_CUDA(cudaMemsetAsync(A.cnt, 0, sizeof(unsigned int) * 4, st));
  // iteration 0: validation, t, e0, loop top; every surviving fit needs its first finite-difference Jacobian
  int cur = 0, rounds = 0;
  span_begin(1);
  if (bucket) {
    const BucketRange R{A.L.off, 0, (int)B};
    int* bhist = (int*)(d + o_bhist);
    int* order = (int*)(d + o_order);
    staged_bucket_hist_kernel<<<nchunks, 256, 0, st>>>(R, bhist, nchunks);
    staged_bucket_scan_kernel<<<1, 1024, 0, st>>>(bhist, kBucketBins * nchunks);
    staged_bucket_scatter_kernel<<<nchunks, 256, 0, st>>>(R, bhist, nchunks, order);
    A.init_order = order;
    c->launches += 3;
  }
  dif_eval_kernel<true><<<grid_for(B, kEvalThreads, c->sm_count * eval_occ), kEvalThreads, eval_smem, st>>>(A, cur);
  span_end();
  c->launches += 1;
  builds(B, 0, cur);
  dif_round_end_kernel<<<1, 1, 0, st>>>(A.cnt, cur, sl.h_count);
  c->launches += 1;
  CPS_CUDA(cudaGetLastError());
  CPS_CUDA(cudaStreamSynchronize(st));
  spans_collect();
  size_t n_solve = sl.h_count[0];
  size_t last_fd = sl.h_count[1], last_upd = sl.h_count[2];
  long long n_builds = (long long)last_fd;
  const bool trace = std::getenv("CPS_DIF_TRACE") != nullptr;
  double tr[3] = {c->prof_ms[0], c->prof_ms[1], c->prof_ms[2]};
  while (n_solve > 0) {
    const int nxt = cur ^ 1;
    CPS_CUDA(cudaMemsetAsync(A.cnt + 0, 0, sizeof(unsigned int) * 2, st));
    CPS_CUDA(cudaMemsetAsync(A.cnt + 2 + nxt, 0, sizeof(unsigned int), st));
    span_begin(0);
    dif_solve_kernel<<<grid_for(n_solve, kSolveThreads, c->sm_count * solve_occ), kSolveThreads, solve_smem, st>>>(A, cur);
    span_end();
    span_begin(1);
    dif_eval_kernel<false><<<grid_for(n_solve, kEvalThreads, c->sm_count * eval_occ), kEvalThreads, eval_smem, st>>>(A, nxt);
    span_end();
    c->launches += 2;
    // the kernels read the real list lengths on the device; the grids follow an upper bound (a block without work
    // leaves at once): at most every solved fit needs a build
    (void)last_upd;
    builds(std::min(n_solve, std::max<size_t>(2 * last_fd + 4096, n_solve / 8)), n_solve, nxt);
    dif_round_end_kernel<<<1, 1, 0, st>>>(A.cnt, nxt, sl.h_count);
    c->launches += 1;
    CPS_CUDA(cudaGetLastError());
    CPS_CUDA(cudaStreamSynchronize(st));
    spans_collect();
    n_solve = sl.h_count[0];
    last_fd = sl.h_count[1];
    last_upd = sl.h_count[2];
    n_builds += (long long)(last_fd + last_upd);
    cur = nxt;
    ++rounds;
    if (trace)
      std::fprintf(stderr, "dif round %d: fd %zu upd %zu -> solve %zu; ms solve %.3f eval %.3f build %.3f\n", rounds, last_fd,
                   last_upd, n_solve, c->prof_ms[0] - tr[0], c->prof_ms[1] - tr[1], c->prof_ms[2] - tr[2]);
    for (int q = 0; q < 3; ++q) tr[q] = c->prof_ms[q];
  }
  c->last_grid = build_grid_max;
  c->last_block = kDifThreads;
  c->last_smem = upd_smem;
  c->last_rounds = rounds;
  if (c->profiling) c->prof_jac_evals += n_builds;
  return CPS_OK;
}

int launch_staged_dif19(cps_lm_ctx* c, cps_lm_slot& sl, cps::LmLaunch& L, cudaStream_t st) {
  const size_t per_fit = sizeof(double) * (size_t)L.n_max * cps::kDifRow;
  const size_t max_fits = std::max<size_t>(1024, kDifMaxJBytes / std::max<size_t>(per_fit, 1));
  int rounds = 0;
  for (size_t first = 0; first < (size_t)L.B; first += max_fits) {
    cps::LmLaunch R = L;
    R.B = (int)std::min(max_fits, (size_t)L.B - first);
    R.p = L.p + first * 19;
    R.off = L.off + first;
    R.counts = L.counts + first * 4;
    R.info = L.info + first * 10;
    R.ret = L.ret + first;
    R.extra = L.extra ? L.extra + first * 2 : nullptr;
    const int rc = launch_staged_dif19_range(c, sl, R, st);
    if (rc != CPS_OK) return rc;
    rounds += c->last_rounds;
  }
  c->last_rounds = rounds;
  return CPS_OK;
}

// dlevmar_der / dlevmar_dif for a large batch of per-contour m = 8 fits: one kernel, a thread per fit
// (cps_lm_image8.cuh).  Work area: [dif: the secant Jacobians, n_max x 8 per resident thread][t of every fit unless
// the caller supplies it]
int launch_image8(cps_lm_ctx* c, cps_lm_slot& sl, cps::LmLaunch& L, cudaStream_t st, bool dif) {
  using namespace cps;
  const int smem = kI8LuDoubles * kI8Threads * (int)sizeof(double) + (kI8Threads / 32) * (int)sizeof(EvalTiles);
  const void* fn = dif ? (const void*)image8_dif_kernel : (const void*)image8_der_kernel;
  CPS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int occ = 0;
  if (dif) CPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, image8_dif_kernel, kI8Threads, smem));
  else CPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, image8_der_kernel, kI8Threads, smem));
  if (occ < 1) {
    cps::set_last_error("image8 kernel does not fit on this device");
    return CPS_ERR_CUDA;
  }
  const long long ctas = ((long long)L.B + kI8Threads - 1) / kI8Threads;
  const int grid = (int)std::max<long long>(1, std::min<long long>(ctas, (long long)c->sm_count * occ));
  const size_t jac_bytes = dif ? sizeof(double) * 8 * (size_t)L.n_max * (size_t)grid * kI8Threads : 0;
  const size_t need = jac_bytes + (L.t ? 0 : sizeof(double) * (size_t)(L.n_max / 2) * (size_t)L.B);
  if (need > sl.staged_bytes) {
    CPS_CUDA(cudaStreamSynchronize(st));
    CPS_CUDA(cudaStreamSynchronize(sl.stream));
    cudaFree(sl.d_staged);
    sl.d_staged = nullptr;
    sl.staged_bytes = 0;
    CPS_CUDA(cudaMalloc(&sl.d_staged, need));
    sl.staged_bytes = need;
  }
  double* tws = (double*)((char*)sl.d_staged + jac_bytes);
  if (dif) image8_dif_kernel<<<grid, kI8Threads, smem, st>>>(L, tws, (double*)sl.d_staged);
  else image8_der_kernel<<<grid, kI8Threads, smem, st>>>(L, tws);
  CPS_CUDA(cudaGetLastError());
  c->last_grid = grid;
  c->last_block = kI8Threads;
  c->last_smem = smem;
  c->launches += 1;
  return CPS_OK;
}

int fill_opts(cps::LmLaunch& L, const double* opts, int variant) {
  // lm_core.c:135-141 / :510-517; defaults levmar.h:87-93
  if (opts) {
    L.tau = opts[0]; L.eps1 = opts[1]; L.eps2 = opts[2]; L.eps2_sq = opts[2] * opts[2]; L.eps3 = opts[3];
    L.delta = 1E-06; L.forward = 1;
    if (variant == CPS_LM_DIF) {
      L.delta = opts[4];
      if (L.delta < 0.0) { L.delta = -L.delta; L.forward = 0; }
    }
  } else {
    L.tau = 1E-03; L.eps1 = 1E-17; L.eps2 = 1E-17; L.eps2_sq = 1E-17 * 1E-17; L.eps3 = 1E-17;
    L.delta = 1E-06; L.forward = 1;
  }
  return CPS_OK;
}

}  // namespace

namespace {
int fit_dev(cps_lm_ctx* c, cps_lm_slot& sl, int model, int variant, int B, double* d_p, const double* d_meas,
            const double* d_t, const long long* d_off, const int* d_counts, int n_max, int itmax,
            const double* opts, double* d_info, int* d_ret, int* d_extra, cudaStream_t st, int path) {
  cps::LmLaunch L;
  std::memset(&L, 0, sizeof L);
  L.B = B;
  L.n_max = (n_max + 3) & ~3;
  L.itmax = itmax;
  fill_opts(L, opts, variant);
  L.p = d_p; L.meas = d_meas; L.t = d_t; L.off = d_off; L.counts = d_counts;
  L.info = d_info; L.ret = d_ret; L.extra = d_extra;
  if (path == CPS_LM_PATH_AUTO) path = c->path;
  // fits longer than the warp-per-fit kernel's compact item list (40 row blocks = 1 280 measurements) always take the
  // staged kernels, which walk the rows of a fit without any such list (the reference accepts 4 x 10 000 points)
  if (model == CPS_MODEL_STEREO19 && variant == CPS_LM_DER && path != CPS_LM_PATH_MONOLITHIC &&
      (path == CPS_LM_PATH_STAGED || B >= kStagedMinFits || L.n_max > 40 * cps::kTileRows))
    return launch_staged_der19(c, sl, L, st, {StagedPiece{0, B, nullptr}});
  // central differences (opts[4] < 0) and fits longer than the staged kernels' row index stay on the warp-per-fit kernel
  if (model == CPS_MODEL_STEREO19 && variant == CPS_LM_DIF && L.forward && path != CPS_LM_PATH_MONOLITHIC &&
      (path == CPS_LM_PATH_STAGED || B >= kDifStagedMinFits))
    return launch_staged_dif19(c, sl, L, st);
  if (model == CPS_MODEL_IMAGE8 && path != CPS_LM_PATH_MONOLITHIC &&
      (path == CPS_LM_PATH_STAGED || B >= kImage8ThreadMinFits))
    return launch_image8(c, sl, L, st, variant == CPS_LM_DIF);
  if (model == CPS_MODEL_STEREO19)
    return variant == CPS_LM_DER ? launch_fit<19, false>(c, sl, L, st) : launch_fit<19, true>(c, sl, L, st);
  return variant == CPS_LM_DER ? launch_fit<8, false>(c, sl, L, st) : launch_fit<8, true>(c, sl, L, st);
}
}  // namespace

namespace {
// the worker of the asynchronous device entry: one call at a time, in the order they were made
void async_worker(cps_lm_ctx* c) {
  cudaSetDevice(c->device);
  for (;;) {
    cps_lm_ctx::AsyncJob j;
    {
      std::unique_lock<std::mutex> lk(c->mu);
      c->cv_job.wait(lk, [&] { return c->stop || !c->jobs.empty(); });
      if (c->jobs.empty()) return;  // stop requested and nothing left
      j = c->jobs.front();
      c->jobs.pop_front();
    }
    cudaStream_t st = c->async_stream;
    int rc = CPS_OK;
    if (cudaStreamWaitEvent(st, j.ev_in, 0) != cudaSuccess) rc = CPS_ERR_CUDA;  // the caller's earlier work on its stream
    cudaEventDestroy(j.ev_in);
    if (rc == CPS_OK)
      rc = fit_dev(c, c->slot[0], j.model, j.variant, j.B, j.d_p, j.d_meas, j.d_t, j.d_off, j.d_counts, j.n_max, j.itmax,
                   j.has_opts ? j.opts : nullptr, j.d_info, j.d_ret, j.d_extra, st, CPS_LM_PATH_AUTO);
    if (cudaStreamSynchronize(st) != cudaSuccess && rc == CPS_OK) rc = CPS_ERR_CUDA;
    __sync_synchronize();
    *c->flag_host = j.seq;  // releases the caller's stream (also after an error: it must not wait for ever)
    {
      std::lock_guard<std::mutex> lk(c->mu);
      c->seq_done = j.seq;
      if (rc != CPS_OK && c->async_rc == CPS_OK) {
        c->async_rc = rc;
        c->async_error = cps_last_error();
      }
    }
    c->cv_done.notify_all();
  }
}
}  // namespace

extern "C" int cps_lm_set_async(cps_lm_ctx* c, int on) {
  if (!c) return CPS_ERR_INVALID;
  if (!on) {
    const int rc = cps_lm_sync(c);
    c->async_on = 0;
    return rc;
  }
  if (c->async_on) return CPS_OK;
  CPS_CUDA(cudaSetDevice(c->device));
  if (!c->wait_value32) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &c->wait_value32, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || !c->wait_value32) {
      c->wait_value32 = nullptr;
      cps::set_last_error("cuStreamWaitValue32 is not available from this driver");
      return CPS_ERR_UNSUPPORTED;
    }
  }
  if (!c->flag_host) {
    unsigned int* h = nullptr;
    CPS_CUDA(cudaHostAlloc((void**)&h, sizeof(unsigned int), cudaHostAllocMapped));
    *h = 0;
    c->flag_host = h;
    void* dptr = nullptr;
    CPS_CUDA(cudaHostGetDevicePointer(&dptr, h, 0));
    c->flag_dev = (CUdeviceptr)dptr;
  }
  if (!c->async_stream) CPS_CUDA(cudaStreamCreateWithFlags(&c->async_stream, cudaStreamNonBlocking));
  if (!c->worker.joinable()) c->worker = std::thread(async_worker, c);
  c->async_on = 1;
  return CPS_OK;
}

extern "C" int cps_lm_fit_batched_dev(cps_lm_ctx* c, int model, int variant, int B, double* d_p,
                                      const double* d_meas, const double* d_t, const long long* d_off,
                                      const int* d_counts, int n_max, int itmax, const double* opts,
                                      double* d_info, int* d_ret, int* d_extra, void* stream) {
  if (!c || B < 0 || !d_p || !d_meas || !d_off || !d_counts || !d_info || !d_ret || n_max < 2 || itmax < 0)
    return CPS_ERR_INVALID;
  if (model != CPS_MODEL_STEREO19 && model != CPS_MODEL_IMAGE8) return CPS_ERR_UNSUPPORTED;
  if (variant != CPS_LM_DER && variant != CPS_LM_DIF) return CPS_ERR_INVALID;
  if (B == 0) return CPS_OK;
  CPS_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;  // NULL = the CUDA default stream, as everywhere in CUDA
  if (c->async_on) {
    // the call is queued for the worker and returns; `stream` is made to wait, on the device, until the worker has
    // finished it, so whatever the caller enqueues next sees the results
    cps_lm_ctx::AsyncJob j;
    j.model = model; j.variant = variant; j.B = B; j.n_max = n_max; j.itmax = itmax;
    j.d_p = d_p; j.d_meas = d_meas; j.d_t = d_t; j.d_off = d_off; j.d_counts = d_counts;
    j.has_opts = opts != nullptr;
    if (opts) std::memcpy(j.opts, opts, sizeof j.opts);
    j.d_info = d_info; j.d_ret = d_ret; j.d_extra = d_extra;
    CPS_CUDA(cudaEventCreateWithFlags(&j.ev_in, cudaEventDisableTiming));
    CPS_CUDA(cudaEventRecord(j.ev_in, st));
    {
      std::lock_guard<std::mutex> lk(c->mu);
      j.seq = ++c->seq_issued;
      c->jobs.push_back(j);
    }
    c->cv_job.notify_one();
    typedef CUresult (*wait_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    if (((wait_fn)c->wait_value32)((CUstream)st, c->flag_dev, j.seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS) {
      cps_lm_sync(c);  // could not make the stream wait: fall back to waiting here
      cps::set_last_error("cuStreamWaitValue32 failed");
      return CPS_ERR_CUDA;
    }
    return CPS_OK;
  }
  // the context has ONE fit queue and work area: a call on another stream queues behind the previous call's kernels
  // instead of resetting the queue they are still reading
  if (!c->ev_last) CPS_CUDA(cudaEventCreateWithFlags(&c->ev_last, cudaEventDisableTiming));
  else CPS_CUDA(cudaStreamWaitEvent(st, c->ev_last, 0));
  const int rc = fit_dev(c, c->slot[0], model, variant, B, d_p, d_meas, d_t, d_off, d_counts, n_max, itmax, opts, d_info,
                         d_ret, d_extra, st, CPS_LM_PATH_AUTO);
  cudaEventRecord(c->ev_last, st);
  return rc;
}

namespace {

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Samples that arrive as integer pixel coordinates (the reference's CvPoint contours; 4 or 2 bytes per coordinate on
// the bus instead of 8) become the doubles the kernels read -- exactly, every pixel coordinate is a double.
template <typename T>
__global__ void pix_to_meas_kernel(const T* __restrict__ src, double* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (double)src[i];
}
inline size_t pix_size(int pix_type) { return pix_type == CPS_PTS_INT16 ? sizeof(short) : sizeof(int); }
// host samples [first, first + count) -> device doubles at `dst`, through `stage` when they are integers (pix_type < 0: doubles)
int copy_samples(const void* host, int pix_type, long long first, long long count, double* dst, void* stage,
                 cudaStream_t st) {
  if (count <= 0) return CPS_OK;
  if (pix_type < 0) {
    CPS_CUDA(cudaMemcpyAsync(dst, (const double*)host + first, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, st));
    return CPS_OK;
  }
  const size_t es = pix_size(pix_type);
  CPS_CUDA(cudaMemcpyAsync(stage, (const char*)host + es * (size_t)first, es * (size_t)count, cudaMemcpyHostToDevice, st));
  const int grid = (int)std::min<long long>((count + 1023) / 1024, 4096);
  if (pix_type == CPS_PTS_INT16) pix_to_meas_kernel<short><<<grid, 256, 0, st>>>((const short*)stage, dst, count);
  else pix_to_meas_kernel<int><<<grid, 256, 0, st>>>((const int*)stage, dst, count);
  CPS_CUDA(cudaGetLastError());
  return CPS_OK;
}

// Host buffers, staged shape: the whole batch gets one staging block; p / offsets / counts go first, then the samples
// in pieces on the copy stream (slot 1), each followed by an event.  The rounds (slot 0's stream) pick a piece up as
// soon as its event has fired, so the fits of the first pieces are already iterating while the rest of the batch
// is still crossing PCIe; results come back in one copy at the end.
int fit_batched_host_staged(cps_lm_ctx* c, int B, double* p, const void* meas, int pix_type, const double* t,
                            const long long* off, const int* counts, long long n_max, int itmax, const double* opts,
                            double* info, int* ret, int* extra) {
  constexpr int m = 19, nseg = 4;
  cps_lm_slot& sl = c->slot[0];
  cudaStream_t st = sl.stream, cp = c->slot[1].stream;
  const long long base = off[0], total = off[B] - base;
  size_t o_p = 0, o_meas = align256(o_p + sizeof(double) * (size_t)B * m);
  size_t o_t = align256(o_meas + sizeof(double) * (size_t)total);
  size_t o_off = align256(o_t + (t ? sizeof(double) * (size_t)(total / 2 + 1) : 0));
  size_t o_cnt = align256(o_off + sizeof(long long) * (size_t)(B + 1));
  size_t o_info = align256(o_cnt + sizeof(int) * (size_t)B * nseg);
  size_t o_ret = align256(o_info + sizeof(double) * (size_t)B * 10);
  size_t o_ext = align256(o_ret + sizeof(int) * (size_t)B);
  size_t o_pix = align256(o_ext + sizeof(int) * (size_t)B * 2);
  const size_t bytes = align256(o_pix + (pix_type < 0 ? 0 : pix_size(pix_type) * (size_t)total));
  CPS_CUDA(cudaStreamSynchronize(st));
  CPS_CUDA(cudaStreamSynchronize(cp));
  if (bytes > sl.stage_bytes) {
    cudaFree(sl.d_stage);
    sl.d_stage = nullptr;
    sl.stage_bytes = 0;
    CPS_CUDA(cudaMalloc(&sl.d_stage, bytes));
    sl.stage_bytes = bytes;
  }
  char* d = sl.d_stage;
  sl.off0.resize((size_t)B + 1);
  for (int f = 0; f <= B; ++f) sl.off0[f] = off[f] - base;
  // piece sizes: small at first so that the rounds can start early, then growing to B/16
  int piece_cap = std::max(8192, (B + 15) / 16);
  if (const char* e = std::getenv("CPS_LM_STAGED_PIECE")) {  // experiment knob
    const int v = std::atoi(e);
    if (v >= 1024) piece_cap = v;
  }
  std::vector<int> bounds(1, 0);
  for (int sz = std::min(piece_cap, std::max(2048, B / 64)); bounds.back() < B; sz = std::min(piece_cap, 2 * sz))
    bounds.push_back(std::min(B, bounds.back() + sz));
  const int n_pieces = (int)bounds.size() - 1;
  std::vector<StagedPiece> pieces((size_t)n_pieces);
  std::vector<cudaEvent_t> events((size_t)n_pieces, nullptr);
  auto destroy_events = [&]() {
    for (cudaEvent_t e : events)
      if (e) cudaEventDestroy(e);
  };
  auto run = [&]() -> int {
    // the rebased offsets live in pageable memory (a copy from there blocks the host until the stream reaches it), so
    // they go first and in one piece
    CPS_CUDA(cudaMemcpyAsync(d + o_off, sl.off0.data(), sizeof(long long) * (size_t)(B + 1), cudaMemcpyHostToDevice, cp));
    for (int k = 0; k < n_pieces; ++k) {
      const int f0 = bounds[k], f1 = bounds[k + 1];
      const long long m0 = sl.off0[f0], m1 = sl.off0[f1];
      // the piece's parameters and segment counts, then its samples
      CPS_CUDA(cudaMemcpyAsync(d + o_p + sizeof(double) * (size_t)f0 * m, p + (size_t)f0 * m,
                               sizeof(double) * (size_t)(f1 - f0) * m, cudaMemcpyHostToDevice, cp));
      CPS_CUDA(cudaMemcpyAsync(d + o_cnt + sizeof(int) * (size_t)f0 * nseg, counts + (size_t)f0 * nseg,
                               sizeof(int) * (size_t)(f1 - f0) * nseg, cudaMemcpyHostToDevice, cp));
      if (m1 > m0) {
        const int rc_s = copy_samples(meas, pix_type, base + m0, m1 - m0, (double*)(d + o_meas) + m0,
                                      d + o_pix + (pix_type < 0 ? 0 : pix_size(pix_type) * (size_t)m0), cp);
        if (rc_s != CPS_OK) return rc_s;
        if (t)
          CPS_CUDA(cudaMemcpyAsync(d + o_t + sizeof(double) * (size_t)(m0 / 2), t + (base + m0) / 2,
                                   sizeof(double) * (size_t)((m1 - m0) / 2), cudaMemcpyHostToDevice, cp));
      }
      CPS_CUDA(cudaEventCreateWithFlags(&events[k], cudaEventDisableTiming));
      CPS_CUDA(cudaEventRecord(events[k], cp));
      pieces[k] = StagedPiece{f0, f1 - f0, events[k]};
    }
    cps::LmLaunch L;
    std::memset(&L, 0, sizeof L);
    L.B = B;
    L.n_max = ((int)n_max + 3) & ~3;
    L.itmax = itmax;
    fill_opts(L, opts, CPS_LM_DER);
    L.p = (double*)(d + o_p);
    L.meas = (const double*)(d + o_meas);
    L.t = t ? (const double*)(d + o_t) : nullptr;
    L.off = (const long long*)(d + o_off);
    L.counts = (const int*)(d + o_cnt);
    L.info = (double*)(d + o_info);
    L.ret = (int*)(d + o_ret);
    L.extra = extra ? (int*)(d + o_ext) : nullptr;
    const int rc = launch_staged_der19(c, sl, L, st, pieces);  // c->bucket_hint: set by fit_batched_host on ragged offsets
    if (rc != CPS_OK) return rc;
    CPS_CUDA(cudaMemcpyAsync(p, d + o_p, sizeof(double) * (size_t)B * m, cudaMemcpyDeviceToHost, st));
    CPS_CUDA(cudaMemcpyAsync(info, d + o_info, sizeof(double) * (size_t)B * 10, cudaMemcpyDeviceToHost, st));
    CPS_CUDA(cudaMemcpyAsync(ret, d + o_ret, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, st));
    if (extra) CPS_CUDA(cudaMemcpyAsync(extra, d + o_ext, sizeof(int) * (size_t)B * 2, cudaMemcpyDeviceToHost, st));
    CPS_CUDA(cudaStreamSynchronize(st));
    return CPS_OK;
  };
  const int rc = run();
  cudaStreamSynchronize(cp);
  cudaStreamSynchronize(st);
  destroy_events();
  return rc;
}

// Host buffers: the batch is cut into chunks of kChunkFits fits; chunk i+1 is copied (and its results of
// chunk i-1 copied back) on the other slot's stream while chunk i computes.  Pinned host memory makes the
// copies truly asynchronous; pageable memory works but serialises.
int fit_batched_host(cps_lm_ctx* c, int model, int variant, int B, double* p, const void* meas, int pix_type,
                     const double* t, const long long* off, const int* counts, int itmax, const double* opts, double* info,
                     int* ret, int* extra) {
  if (!c || B < 0 || !p || !meas || !off || !counts || !info || !ret || itmax < 0) return CPS_ERR_INVALID;
  if (model != CPS_MODEL_STEREO19 && model != CPS_MODEL_IMAGE8) return CPS_ERR_UNSUPPORTED;
  if (B == 0) return CPS_OK;
  const int m = (model == CPS_MODEL_STEREO19) ? 19 : 8;
  const int nseg = (model == CPS_MODEL_STEREO19) ? 4 : 1;
  long long n_max = 2, n_min = (1LL << 40);
  if (off[0] < 0) return CPS_ERR_INVALID;
  for (int f = 0; f < B; ++f) {
    const long long n = off[f + 1] - off[f];
    if (n < 0) return CPS_ERR_INVALID;
    n_max = std::max(n_max, n);
    n_min = std::min(n_min, n);
  }
  if (n_max > (1 << 24)) return CPS_ERR_INVALID;
  CPS_CUDA(cudaSetDevice(c->device));
  // ragged offsets (the longest fit at least two 32-row blocks longer than the shortest): the staged shapes order the
  // fits by length before they enter their lists
  c->bucket_hint = ((n_max + 31) / 32 - (n_min + 31) / 32 >= 2) ? 1 : 0;
  struct HintReset { cps_lm_ctx* c; ~HintReset() { c->bucket_hint = 0; } } hint_reset{c};
  const bool staged = model == CPS_MODEL_STEREO19 && variant == CPS_LM_DER && c->path != CPS_LM_PATH_MONOLITHIC &&
                      (c->path == CPS_LM_PATH_STAGED || B >= kStagedMinFits || n_max > 40 * cps::kTileRows);
  if (staged) return fit_batched_host_staged(c, B, p, meas, pix_type, t, off, counts, n_max, itmax, opts, info, ret, extra);
  // staged dif: larger chunks (the late rounds of a chunk run at the latency floor of the kernels)
  const bool dif_staged = model == CPS_MODEL_STEREO19 && variant == CPS_LM_DIF && (!opts || opts[4] >= 0.0) &&
                          c->path != CPS_LM_PATH_MONOLITHIC && (c->path == CPS_LM_PATH_STAGED || B >= kDifStagedMinFits);
  const int chunk = dif_staged ? kDifChunkFits : kChunkFits;
  const int n_chunks = (B + chunk - 1) / chunk;
  struct Staged {
    int f0, nb;
    size_t o_p, o_meas, o_t, o_off, o_cnt, o_info, o_ret, o_ext, o_pix;
  } desc[kSlots];
  // copies of chunk ci into its slot's staging block: p | meas | t | off | counts | info | ret | extra
  auto stage_in = [&](int ci) -> int {
    cps_lm_slot& sl = c->slot[ci % kSlots];
    Staged& D = desc[ci % kSlots];
    cudaStream_t st = sl.stream;
    D.f0 = ci * chunk;
    D.nb = std::min(chunk, B - D.f0);
    const int f0 = D.f0, nb = D.nb;
    const long long base = off[f0], total = off[f0 + nb] - base;
    D.o_p = 0;
    D.o_meas = align256(D.o_p + sizeof(double) * (size_t)nb * m);
    D.o_t = align256(D.o_meas + sizeof(double) * (size_t)total);
    D.o_off = align256(D.o_t + (t ? sizeof(double) * (size_t)(total / 2 + 1) : 0));
    D.o_cnt = align256(D.o_off + sizeof(long long) * (size_t)(nb + 1));
    D.o_info = align256(D.o_cnt + sizeof(int) * (size_t)nb * nseg);
    D.o_ret = align256(D.o_info + sizeof(double) * (size_t)nb * 10);
    D.o_ext = align256(D.o_ret + sizeof(int) * (size_t)nb);
    D.o_pix = align256(D.o_ext + sizeof(int) * (size_t)nb * 2);
    const size_t bytes = align256(D.o_pix + (pix_type < 0 ? 0 : pix_size(pix_type) * (size_t)total));
    // the slot's previous chunk must have drained before its staging block / offsets are reused
    CPS_CUDA(cudaStreamSynchronize(st));
    if (bytes > sl.stage_bytes) {
      cudaFree(sl.d_stage);
      sl.d_stage = nullptr;
      sl.stage_bytes = 0;
      CPS_CUDA(cudaMalloc(&sl.d_stage, bytes));
      sl.stage_bytes = bytes;
    }
    char* d = sl.d_stage;
    sl.off0.resize((size_t)nb + 1);
    for (int f = 0; f <= nb; ++f) sl.off0[f] = off[f0 + f] - base;
    CPS_CUDA(cudaMemcpyAsync(d + D.o_p, p + (size_t)f0 * m, sizeof(double) * (size_t)nb * m, cudaMemcpyHostToDevice, st));
    {
      const int rc_s = copy_samples(meas, pix_type, base, total, (double*)(d + D.o_meas), d + D.o_pix, st);
      if (rc_s != CPS_OK) return rc_s;
    }
    if (t)
      CPS_CUDA(cudaMemcpyAsync(d + D.o_t, t + base / 2, sizeof(double) * (size_t)(total / 2), cudaMemcpyHostToDevice, st));
    CPS_CUDA(cudaMemcpyAsync(d + D.o_off, sl.off0.data(), sizeof(long long) * (size_t)(nb + 1), cudaMemcpyHostToDevice, st));
    CPS_CUDA(cudaMemcpyAsync(d + D.o_cnt, counts + (size_t)f0 * nseg, sizeof(int) * (size_t)nb * nseg,
                             cudaMemcpyHostToDevice, st));
    return CPS_OK;
  };
  // kernels of chunk ci (the staged shape blocks the host while its rounds run: the other slot's copies proceed
  // meanwhile) and the copies back
  auto run = [&](int ci) -> int {
    cps_lm_slot& sl = c->slot[ci % kSlots];
    const Staged& D = desc[ci % kSlots];
    cudaStream_t st = sl.stream;
    char* d = sl.d_stage;
    const int f0 = D.f0, nb = D.nb;
    int rc = fit_dev(c, sl, model, variant, nb, (double*)(d + D.o_p), (const double*)(d + D.o_meas),
                     t ? (const double*)(d + D.o_t) : nullptr, (const long long*)(d + D.o_off),
                     (const int*)(d + D.o_cnt), (int)n_max, itmax, opts, (double*)(d + D.o_info), (int*)(d + D.o_ret),
                     extra ? (int*)(d + D.o_ext) : nullptr, st,
                     // a chunk of m = 19 fits stays on the one-launch kernel (the staged rounds would block this
                     // pipeline); the m = 8 thread-per-fit kernels are single launches and follow the context's choice
                     (model == CPS_MODEL_IMAGE8 || dif_staged) ? CPS_LM_PATH_AUTO : CPS_LM_PATH_MONOLITHIC);
    if (rc != CPS_OK) return rc;
    CPS_CUDA(cudaMemcpyAsync(p + (size_t)f0 * m, d + D.o_p, sizeof(double) * (size_t)nb * m, cudaMemcpyDeviceToHost, st));
    CPS_CUDA(cudaMemcpyAsync(info + (size_t)f0 * 10, d + D.o_info, sizeof(double) * (size_t)nb * 10,
                             cudaMemcpyDeviceToHost, st));
    CPS_CUDA(cudaMemcpyAsync(ret + f0, d + D.o_ret, sizeof(int) * (size_t)nb, cudaMemcpyDeviceToHost, st));
    if (extra)
      CPS_CUDA(cudaMemcpyAsync(extra + (size_t)f0 * 2, d + D.o_ext, sizeof(int) * (size_t)nb * 2, cudaMemcpyDeviceToHost, st));
    return CPS_OK;
  };
  int rc = stage_in(0);
  for (int ci = 0; ci < n_chunks && rc == CPS_OK; ++ci) {
    if (ci + 1 < n_chunks) rc = stage_in(ci + 1);
    if (rc == CPS_OK) rc = run(ci);
  }
  if (rc != CPS_OK) {
    // copies into / out of the caller's buffers may still be queued: they must not outlive this call
    for (int i = 0; i < kSlots; ++i) cudaStreamSynchronize(c->slot[i].stream);
    return rc;
  }
  for (int i = 0; i < kSlots; ++i) CPS_CUDA(cudaStreamSynchronize(c->slot[i].stream));
  return CPS_OK;
}

}  // namespace

extern "C" int cps_dlevmar_der_batched(cps_lm_ctx* c, int model, int B, double* p, const double* meas,
                                       const double* t, const long long* off, const int* counts, int itmax,
                                       const double* opts, double* info, int* ret, int* extra) {
  return fit_batched_host(c, model, CPS_LM_DER, B, p, meas, -1, t, off, counts, itmax, opts, info, ret, extra);
}

extern "C" int cps_dlevmar_dif_batched(cps_lm_ctx* c, int model, int B, double* p, const double* meas,
                                       const double* t, const long long* off, const int* counts, int itmax,
                                       const double* opts, double* info, int* ret, int* extra) {
  return fit_batched_host(c, model, CPS_LM_DIF, B, p, meas, -1, t, off, counts, itmax, opts, info, ret, extra);
}

extern "C" int cps_dlevmar_batched_pix(cps_lm_ctx* c, int model, int variant, int B, double* p, const void* pix, int pix_type,
                                       const long long* off, const int* counts, int itmax, const double* opts,
                                       double* info, int* ret, int* extra) {
  if (pix_type != CPS_PTS_INT32 && pix_type != CPS_PTS_INT16) return CPS_ERR_INVALID;
  if (variant != CPS_LM_DER && variant != CPS_LM_DIF) return CPS_ERR_INVALID;
  return fit_batched_host(c, model, variant, B, p, pix, pix_type, nullptr, off, counts, itmax, opts, info, ret, extra);
}

// ---- fit_curve's initial guess -------------------------------------------------------------------------------
extern "C" int cps_fit_curve_init_batched_dev(cps_lm_ctx* c, int B, const double* d_meas, const long long* d_off,
                                              const int* d_counts, double* d_p, void* stream) {
  if (!c || B < 0 || (B && (!d_meas || !d_off || !d_counts || !d_p))) return CPS_ERR_INVALID;
  if (B == 0) return CPS_OK;
  CPS_CUDA(cudaSetDevice(c->device));
  cps::lm_init_guess_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, d_meas, d_off, d_counts, d_p);
  CPS_CUDA(cudaGetLastError());
  c->launches += 1;
  return CPS_OK;
}

extern "C" int cps_fit_curve_init_batched(cps_lm_ctx* c, int B, const double* meas, const long long* off,
                                          const int* counts, double* p_out) {
  if (!c || B < 0 || (B && (!meas || !off || !counts || !p_out))) return CPS_ERR_INVALID;
  if (B == 0) return CPS_OK;
  if (off[0] < 0 || off[B] < off[0]) return CPS_ERR_INVALID;
  CPS_CUDA(cudaSetDevice(c->device));
  cps_lm_slot& sl = c->slot[0];
  cudaStream_t st = sl.stream;
  const long long total = off[B] - off[0];
  size_t o_meas = 0, o_off = align256(sizeof(double) * (size_t)total);
  size_t o_cnt = align256(o_off + sizeof(long long) * (size_t)(B + 1));
  size_t o_p = align256(o_cnt + sizeof(int) * (size_t)B * 4);
  size_t bytes = align256(o_p + sizeof(double) * (size_t)B * 19);
  CPS_CUDA(cudaStreamSynchronize(st));
  if (bytes > sl.stage_bytes) {
    cudaFree(sl.d_stage);
    sl.d_stage = nullptr;
    sl.stage_bytes = 0;
    CPS_CUDA(cudaMalloc(&sl.d_stage, bytes));
    sl.stage_bytes = bytes;
  }
  char* d = sl.d_stage;
  sl.off0.resize((size_t)B + 1);
  for (int f = 0; f <= B; ++f) sl.off0[f] = off[f] - off[0];
  CPS_CUDA(cudaMemcpyAsync(d + o_meas, meas + off[0], sizeof(double) * (size_t)total, cudaMemcpyHostToDevice, st));
  CPS_CUDA(cudaMemcpyAsync(d + o_off, sl.off0.data(), sizeof(long long) * (size_t)(B + 1), cudaMemcpyHostToDevice, st));
  CPS_CUDA(cudaMemcpyAsync(d + o_cnt, counts, sizeof(int) * (size_t)B * 4, cudaMemcpyHostToDevice, st));
  int rc = cps_fit_curve_init_batched_dev(c, B, (const double*)(d + o_meas), (const long long*)(d + o_off),
                                          (const int*)(d + o_cnt), (double*)(d + o_p), (void*)st);
  if (rc != CPS_OK) return rc;
  CPS_CUDA(cudaMemcpyAsync(p_out, d + o_p, sizeof(double) * (size_t)B * 19, cudaMemcpyDeviceToHost, st));
  CPS_CUDA(cudaStreamSynchronize(st));
  return CPS_OK;
}

// ---- cleanup_and_group_edges (curveFittingClass.cpp:1207-1418): contours -> packed fit inputs --------------------
namespace {
constexpr int kLrOffset = 0;     // LEFT_RIGHT_OFFSET, common.h:108
constexpr int kLenCurve = 100;   // LENGTH_EACH_CURVE, common.h:112

// one thread per frame walks the reference's trimming loops (integer compares on image rows, inherently serial
// inside a frame, independent across frames); rng = kept index range per segment, cnt = points kept
template <class PT>
__global__ void __launch_bounds__(128) edges_trim_kernel(int F, const PT* __restrict__ pts,
                                                         const long long* __restrict__ seg_off, int reset_da,
                                                         const double* __restrict__ tracked, int* __restrict__ rng,
                                                         int* __restrict__ cnt, int* __restrict__ valid) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const PT* seg[4];
  int size[4], r0[4], r1[4], ys[4], ye[4];
  bool ok = true;
  for (int s = 0; s < 4; ++s) {
    seg[s] = pts + 2 * seg_off[4 * f + s];
    size[s] = (int)(seg_off[4 * f + s + 1] - seg_off[4 * f + s]);
    r0[s] = 0;
    r1[s] = size[s] - 1;
    if (size[s] < 10) ok = false;
  }
#define YAT(s, k) ((int)seg[s][2 * (k) + 1])
  if (ok) {
    for (int s = 0; s < 4; ++s) { ys[s] = YAT(s, 0); ye[s] = YAT(s, size[s] - 1); }
    for (int i = 0; i < 2 && ok; ++i) {  // align the starts of the left / right view of curve i
      const int L = i, R = i + 2;
      if (ys[L] > ys[R] - kLrOffset) {
        while (ok && ys[L] > ys[R] - kLrOffset) {
          if (++r0[L] >= size[L]) { ok = false; break; }
          ys[L] = YAT(L, r0[L]);
        }
      } else if (ys[L] < ys[R] - kLrOffset) {
        while (ok && ys[L] < ys[R] - kLrOffset) {
          if (++r0[R] >= size[R]) { ok = false; break; }
          ys[R] = YAT(R, r0[R]);
        }
      }
    }
    for (int i = 0; i < 2 && ok; ++i) {
      const int L = i, R = i + 2;
      if (ye[L] < ye[R] - kLrOffset) {
        while (ye[L] < ye[R] - kLrOffset) {
          if (--r1[L] < 0) { ok = false; break; }
          ye[L] = YAT(L, r1[L]);
        }
      } else if (ye[L] > ye[R] - kLrOffset) {
        while (ye[L] > ye[R] - kLrOffset) {
          if (--r1[R] < 0) { ok = false; break; }
          ye[R] = YAT(R, r1[R]);
        }
      }
      while (ok && ye[L] < ys[L] - kLenCurve) {
        if (--r1[L] < 0) { ok = false; break; }
        ye[L] = YAT(L, r1[L]);
      }
      while (ok && ye[R] < ys[R] - kLenCurve) {
        if (--r1[R] < 0) { ok = false; break; }
        ye[R] = YAT(R, r1[R]);
      }
      const double te = tracked[2 * f + i];
      while (ok && ye[L] > te && !reset_da) {
        ++r0[L]; ++r1[L];
        if (r0[L] >= size[L] || r1[L] >= size[L]) { ok = false; break; }
        ye[L] = YAT(L, r1[L]);
      }
      while (ok && ye[R] > te + kLrOffset && !reset_da) {
        ++r1[R]; ++r0[R];
        if (r0[R] >= size[R] || r1[R] >= size[R]) { ok = false; break; }
        ye[R] = YAT(R, r1[R]);
      }
    }
    for (int i = 0; i < 2 && ok; ++i) {  // both curves of an image end on the same row
      const int A = 2 * i, B = 2 * i + 1;
      if (ye[A] > ye[B]) {
        while (ye[A] > ye[B]) {
          ++r0[A]; ++r1[A];
          if (r0[A] >= size[A] || r1[A] >= size[A]) { ok = false; break; }
          ye[A] = YAT(A, r1[A]);
        }
      } else {
        while (ye[A] < ye[B]) {
          ++r0[B]; ++r1[B];
          if (r0[B] >= size[B] || r1[B] >= size[B]) { ok = false; break; }
          ye[B] = YAT(B, r1[B]);
        }
      }
    }
  }
#undef YAT
  for (int s = 0; s < 4; ++s) {
    int c = ok ? min(r1[s] - r0[s] + 1, 10000) : 0;
    if (ok && c < 1) ok = false;
    rng[8 * f + 2 * s] = r0[s];
    rng[8 * f + 2 * s + 1] = r1[s];
    cnt[4 * f + s] = c;
  }
  if (!ok)
    for (int s = 0; s < 4; ++s) cnt[4 * f + s] = 0;
  valid[f] = ok ? 1 : 0;
}

// pack the kept points as doubles in the fit's measurement order [L c1 | L c2 | R c1 | R c2] (fit_curve :435-489;
// the reference copies into four CvMat :1395-1413)
template <class PT>
__global__ void __launch_bounds__(256) edges_pack_kernel(int F, const PT* __restrict__ pts,
                                                         const long long* __restrict__ seg_off,
                                                         const int* __restrict__ rng, const int* __restrict__ cnt,
                                                         const long long* __restrict__ off, double* __restrict__ meas) {
  const int f = blockIdx.x;
  if (f >= F) return;
  int c[4], base[4];
  int acc = 0;
  for (int s = 0; s < 4; ++s) { c[s] = cnt[4 * f + s]; base[s] = acc; acc += c[s]; }
  double* out = meas + off[f];
  for (int k = threadIdx.x; k < acc; k += blockDim.x) {
    const int s = (k >= base[1]) + (k >= base[2]) + (k >= base[3]);
    const int idx = rng[8 * f + 2 * s] + (k - base[s]);
    const PT* p = pts + 2 * (seg_off[4 * f + s] + idx);
    out[2 * k] = (double)p[0];
    out[2 * k + 1] = (double)p[1];
  }
}
}  // namespace

extern "C" int cps_cleanup_and_group_edges_batched(cps_lm_ctx* c, int F, const int* pts, const long long* seg_off,
                                                   int reset_data_assoc, const double* tracked_endpt_y, double* meas_out,
                                                   long long* off_out, int* counts_out, int* valid_out) {
  if (!c || F < 0 || (F && (!pts || !seg_off || !tracked_endpt_y || !meas_out || !off_out || !counts_out || !valid_out)))
    return CPS_ERR_INVALID;
  if (F == 0) { if (off_out) off_out[0] = 0; return CPS_OK; }
  for (int k = 0; k < 4 * F; ++k)
    if (seg_off[k + 1] < seg_off[k]) return CPS_ERR_INVALID;
  CPS_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->slot[0].stream;
  const long long npts = seg_off[4 * F] - seg_off[0];
  int *d_pts = nullptr, *d_rng = nullptr, *d_cnt = nullptr, *d_valid = nullptr;
  long long *d_seg = nullptr, *d_off = nullptr;
  double *d_tr = nullptr, *d_meas = nullptr;
  std::vector<long long> seg0((size_t)4 * F + 1);
  for (int k = 0; k <= 4 * F; ++k) seg0[k] = seg_off[k] - seg_off[0];
  auto cleanup = [&]() {
    cudaFree(d_pts); cudaFree(d_rng); cudaFree(d_cnt); cudaFree(d_valid); cudaFree(d_seg); cudaFree(d_off);
    cudaFree(d_tr); cudaFree(d_meas);
  };
  auto run = [&]() -> int {
    CPS_CUDA(cudaMalloc(&d_pts, sizeof(int) * 2 * (size_t)std::max<long long>(npts, 1)));
    CPS_CUDA(cudaMalloc(&d_seg, sizeof(long long) * ((size_t)4 * F + 1)));
    CPS_CUDA(cudaMalloc(&d_tr, sizeof(double) * 2 * (size_t)F));
    CPS_CUDA(cudaMalloc(&d_rng, sizeof(int) * 8 * (size_t)F));
    CPS_CUDA(cudaMalloc(&d_cnt, sizeof(int) * 4 * (size_t)F));
    CPS_CUDA(cudaMalloc(&d_valid, sizeof(int) * (size_t)F));
    CPS_CUDA(cudaMalloc(&d_off, sizeof(long long) * ((size_t)F + 1)));
    CPS_CUDA(cudaMemcpyAsync(d_pts, pts + 2 * seg_off[0], sizeof(int) * 2 * (size_t)npts, cudaMemcpyHostToDevice, st));
    CPS_CUDA(cudaMemcpyAsync(d_seg, seg0.data(), sizeof(long long) * ((size_t)4 * F + 1), cudaMemcpyHostToDevice, st));
    CPS_CUDA(cudaMemcpyAsync(d_tr, tracked_endpt_y, sizeof(double) * 2 * (size_t)F, cudaMemcpyHostToDevice, st));
    edges_trim_kernel<int><<<(F + 127) / 128, 128, 0, st>>>(F, d_pts, d_seg, reset_data_assoc, d_tr, d_rng, d_cnt, d_valid);
    CPS_CUDA(cudaGetLastError());
    CPS_CUDA(cudaMemcpyAsync(counts_out, d_cnt, sizeof(int) * 4 * (size_t)F, cudaMemcpyDeviceToHost, st));
    CPS_CUDA(cudaMemcpyAsync(valid_out, d_valid, sizeof(int) * (size_t)F, cudaMemcpyDeviceToHost, st));
    CPS_CUDA(cudaStreamSynchronize(st));
    off_out[0] = 0;  // exclusive scan of the per-frame measurement counts (host: F integers)
    for (int f = 0; f < F; ++f)
      off_out[f + 1] = off_out[f] + 2LL * (counts_out[4 * f] + counts_out[4 * f + 1] + counts_out[4 * f + 2] + counts_out[4 * f + 3]);
    const long long total = off_out[F];
    CPS_CUDA(cudaMalloc(&d_meas, sizeof(double) * (size_t)std::max<long long>(total, 1)));
    CPS_CUDA(cudaMemcpyAsync(d_off, off_out, sizeof(long long) * ((size_t)F + 1), cudaMemcpyHostToDevice, st));
    edges_pack_kernel<int><<<F, 256, 0, st>>>(F, d_pts, d_seg, d_rng, d_cnt, d_off, d_meas);
    CPS_CUDA(cudaGetLastError());
    if (total) CPS_CUDA(cudaMemcpyAsync(meas_out, d_meas, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost, st));
    CPS_CUDA(cudaStreamSynchronize(st));
    c->launches += 2;
    return CPS_OK;
  };
  const int rc = run();
  cleanup();
  return rc;
}

// ---- main.cpp:425-445 as ONE call: cleanup_and_group_edges -> fit_curve, nothing returns to the host in between ---
namespace {
// exclusive scan of the per-frame measurement counts (2 * points kept): block-local scan, scan of the block
// totals by one block, offsets added back -- F <= 1024 * 1024 frames per call
constexpr int kScanBlock = 1024;
__global__ void __launch_bounds__(kScanBlock) frames_scan_local(int F, const int* __restrict__ cnt,
                                                                long long* __restrict__ off, long long* __restrict__ totals) {
  __shared__ long long sh[kScanBlock];
  const int f = blockIdx.x * kScanBlock + threadIdx.x;
  long long v = 0;
  if (f < F) v = 2LL * ((long long)cnt[4 * f] + cnt[4 * f + 1] + cnt[4 * f + 2] + cnt[4 * f + 3]);
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int d = 1; d < kScanBlock; d <<= 1) {
    const long long t = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  if (f < F) off[f + 1] = sh[threadIdx.x];  // inclusive, block-local
  if (threadIdx.x == kScanBlock - 1) totals[blockIdx.x] = sh[threadIdx.x];
}
__global__ void __launch_bounds__(kScanBlock) frames_scan_totals(int nb, long long* __restrict__ totals) {
  __shared__ long long sh[kScanBlock];
  long long v = threadIdx.x < nb ? totals[threadIdx.x] : 0;
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int d = 1; d < kScanBlock; d <<= 1) {
    const long long t = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
    __syncthreads();
    sh[threadIdx.x] += t;
    __syncthreads();
  }
  if (threadIdx.x < nb) totals[threadIdx.x] = sh[threadIdx.x] - v;  // exclusive
}
__global__ void __launch_bounds__(kScanBlock) frames_scan_add(int F, long long* __restrict__ off,
                                                              const long long* __restrict__ totals) {
  const int f = blockIdx.x * kScanBlock + threadIdx.x;
  if (f < F) off[f + 1] += totals[blockIdx.x];
  if (f == 0) off[0] = 0;
}

// fit_curve's post-normalisation (curveFittingClass.cpp:684-711) and fitting_error = info[1] (:683); frames the
// trimming rejected are marked CPS_FIT_SKIPPED
__global__ void __launch_bounds__(128) frames_postfit_kernel(int F, const int* __restrict__ valid, double* __restrict__ p,
                                                             double* __restrict__ info, int* __restrict__ ret,
                                                             double* __restrict__ fitting_error) {
  constexpr double kPi = 3.1415926535;  // common.h:114
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  double* q = p + (size_t)f * 19;
  if (!valid[f]) {
    for (int i = 0; i < 19; ++i) q[i] = 0.0;
    for (int i = 0; i < 10; ++i) info[(size_t)f * 10 + i] = 0.0;
    ret[f] = CPS_FIT_SKIPPED;
    if (fitting_error) fitting_error[f] = 0.0;
    return;
  }
  if (fitting_error) fitting_error[f] = info[(size_t)f * 10 + 1];
  if (ret[f] < 0 && ret[f] != CPS_FIT_LM_ERROR) return;  // refused fit: parameters untouched
  if (q[18] > 0.0) {
    for (int i = 0; i < 8; ++i) q[2 * i + 1] *= -1.0;
    q[16] *= -1.0;
    q[17] += kPi;
    q[18] *= -1.0;
  }
  if (q[16] < -kPi / 2.0) {
    for (int i = 0; i < 8; ++i) q[2 * i + 1] *= -1.0;
    q[17] = (-kPi - q[17]);
    q[18] += kPi;
  }
  if (fabs(q[16]) < 1e9 && fabs(q[17]) < 1e9) {  // bounded: a wild angle cannot spin
    while (q[16] > kPi) q[16] -= 2 * kPi;
    while (q[16] < -kPi) q[16] += 2 * kPi;
    while (q[17] > kPi) q[17] -= 2 * kPi;
    while (q[17] < -kPi) q[17] += 2 * kPi;
  }
}

template <class PT>
int fit_frames_impl(cps_lm_ctx* c, int variant, int F, const PT* pts, const long long* seg_off, int reset_data_assoc,
                    const double* tracked_endpt_y, int itmax, const double* opts, double* p_out, double* info, int* ret,
                    int* valid, double* fitting_error) {
  CPS_CUDA(cudaSetDevice(c->device));
  cps_lm_slot& sl = c->slot[0];
  cudaStream_t st = sl.stream;
  const long long npts = seg_off[4 * F] - seg_off[0];
  long long n_max = 2;
  for (int f = 0; f < F; ++f) {
    long long n = 0;
    for (int s = 0; s < 4; ++s) n += std::min<long long>(seg_off[4 * f + s + 1] - seg_off[4 * f + s], 10000);
    n_max = std::max(n_max, 2 * n);
  }
  if (n_max > (1 << 24) || F > kScanBlock * kScanBlock) return CPS_ERR_INVALID;
  const int nb = (F + kScanBlock - 1) / kScanBlock;
  // one grow-only block: raw points | segment offsets | tracked | ranges | counts | valid | offsets | block totals |
  // packed measurements (at most 2 doubles per raw point) | p | info | ret | fitting_error
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o = align256(o + bytes); return at; };
  const size_t o_pts = take(sizeof(PT) * 2 * (size_t)std::max<long long>(npts, 1));
  const size_t o_seg = take(sizeof(long long) * ((size_t)4 * F + 1));
  const size_t o_tr = take(sizeof(double) * 2 * (size_t)F);
  const size_t o_rng = take(sizeof(int) * 8 * (size_t)F);
  const size_t o_cnt = take(sizeof(int) * 4 * (size_t)F);
  const size_t o_val = take(sizeof(int) * (size_t)F);
  const size_t o_off = take(sizeof(long long) * ((size_t)F + 1));
  const size_t o_tot = take(sizeof(long long) * kScanBlock);
  const size_t o_meas = take(sizeof(double) * 2 * (size_t)std::max<long long>(npts, 1));
  const size_t o_p = take(sizeof(double) * 19 * (size_t)F);
  const size_t o_info = take(sizeof(double) * 10 * (size_t)F);
  const size_t o_ret = take(sizeof(int) * (size_t)F);
  const size_t o_err = take(sizeof(double) * (size_t)F);
  CPS_CUDA(cudaStreamSynchronize(st));  // the block (and off0 below) may still serve the previous call
  if (o > sl.stage_bytes) {
    cudaFree(sl.d_stage);
    sl.d_stage = nullptr;
    sl.stage_bytes = 0;
    CPS_CUDA(cudaMalloc(&sl.d_stage, o));
    sl.stage_bytes = o;
  }
  char* d = sl.d_stage;
  sl.off0.resize((size_t)4 * F + 1);
  for (int k = 0; k <= 4 * F; ++k) sl.off0[k] = seg_off[k] - seg_off[0];
  PT* d_pts = (PT*)(d + o_pts);
  long long* d_seg = (long long*)(d + o_seg);
  double* d_tr = (double*)(d + o_tr);
  int* d_rng = (int*)(d + o_rng);
  int* d_cnt = (int*)(d + o_cnt);
  int* d_val = (int*)(d + o_val);
  long long* d_off = (long long*)(d + o_off);
  long long* d_tot = (long long*)(d + o_tot);
  double* d_meas = (double*)(d + o_meas);
  double* d_p = (double*)(d + o_p);
  double* d_info = (double*)(d + o_info);
  int* d_ret = (int*)(d + o_ret);
  double* d_err = (double*)(d + o_err);
  auto body = [&]() -> int {
    CPS_CUDA(cudaMemcpyAsync(d_pts, pts + 2 * seg_off[0], sizeof(PT) * 2 * (size_t)npts, cudaMemcpyHostToDevice, st));
    CPS_CUDA(cudaMemcpyAsync(d_seg, sl.off0.data(), sizeof(long long) * ((size_t)4 * F + 1), cudaMemcpyHostToDevice, st));
    CPS_CUDA(cudaMemcpyAsync(d_tr, tracked_endpt_y, sizeof(double) * 2 * (size_t)F, cudaMemcpyHostToDevice, st));
    edges_trim_kernel<PT><<<(F + 127) / 128, 128, 0, st>>>(F, d_pts, d_seg, reset_data_assoc, d_tr, d_rng, d_cnt, d_val);
    frames_scan_local<<<nb, kScanBlock, 0, st>>>(F, d_cnt, d_off, d_tot);
    frames_scan_totals<<<1, kScanBlock, 0, st>>>(nb, d_tot);
    frames_scan_add<<<nb, kScanBlock, 0, st>>>(F, d_off, d_tot);
    edges_pack_kernel<PT><<<F, 128, 0, st>>>(F, d_pts, d_seg, d_rng, d_cnt, d_off, d_meas);
    cps::lm_init_guess_kernel<<<(F + 127) / 128, 128, 0, st>>>(F, d_meas, d_off, d_cnt, d_p);
    CPS_CUDA(cudaGetLastError());
    c->launches += 6;
    c->bucket_hint = 1;  // contours of real lengths: the fits enter the staged lists ordered by length
    struct HintReset { cps_lm_ctx* c; ~HintReset() { c->bucket_hint = 0; } } hint_reset{c};
    const int rc = fit_dev(c, sl, CPS_MODEL_STEREO19, variant, F, d_p, d_meas, nullptr, d_off, d_cnt, (int)n_max, itmax, opts,
                           d_info, d_ret, nullptr, st, CPS_LM_PATH_AUTO);
    if (rc != CPS_OK) return rc;
    frames_postfit_kernel<<<(F + 127) / 128, 128, 0, st>>>(F, d_val, d_p, d_info, d_ret, d_err);
    CPS_CUDA(cudaGetLastError());
    c->launches += 1;
    CPS_CUDA(cudaMemcpyAsync(p_out, d_p, sizeof(double) * 19 * (size_t)F, cudaMemcpyDeviceToHost, st));
    CPS_CUDA(cudaMemcpyAsync(info, d_info, sizeof(double) * 10 * (size_t)F, cudaMemcpyDeviceToHost, st));
    CPS_CUDA(cudaMemcpyAsync(ret, d_ret, sizeof(int) * (size_t)F, cudaMemcpyDeviceToHost, st));
    CPS_CUDA(cudaMemcpyAsync(valid, d_val, sizeof(int) * (size_t)F, cudaMemcpyDeviceToHost, st));
    if (fitting_error)
      CPS_CUDA(cudaMemcpyAsync(fitting_error, d_err, sizeof(double) * (size_t)F, cudaMemcpyDeviceToHost, st));
    return CPS_OK;
  };
  const int rc = body();
  const cudaError_t e = cudaStreamSynchronize(st);  // also on failure: no copy may still be touching the caller's buffers
  if (rc != CPS_OK) return rc;
  CPS_CUDA(e);
  return CPS_OK;
}
}  // namespace

extern "C" int cps_fit_frames_batched(cps_lm_ctx* c, int variant, int F, const void* pts, int pts_type,
                                      const long long* seg_off, int reset_data_assoc, const double* tracked_endpt_y,
                                      int itmax, const double* opts, double* p_out, double* info, int* ret, int* valid,
                                      double* fitting_error) {
  if (!c || F < 0 || itmax < 0 || (variant != CPS_LM_DER && variant != CPS_LM_DIF)) return CPS_ERR_INVALID;
  if (F == 0) return CPS_OK;
  if (!pts || !seg_off || !tracked_endpt_y || !p_out || !info || !ret || !valid) return CPS_ERR_INVALID;
  for (int k = 0; k < 4 * F; ++k)
    if (seg_off[k + 1] < seg_off[k]) return CPS_ERR_INVALID;
  if (pts_type == CPS_PTS_INT32)
    return fit_frames_impl<int>(c, variant, F, (const int*)pts, seg_off, reset_data_assoc, tracked_endpt_y, itmax, opts,
                                p_out, info, ret, valid, fitting_error);
  if (pts_type == CPS_PTS_INT16)
    return fit_frames_impl<short>(c, variant, F, (const short*)pts, seg_off, reset_data_assoc, tracked_endpt_y, itmax, opts,
                                  p_out, info, ret, valid, fitting_error);
  return CPS_ERR_INVALID;
}

// ---- same-signature shims ---------------------------------------------------------------------------
// The callbacks are tokens: the device model is selected by comparing pointers; calling them is an error.
extern "C" void cps_modelfunc(double*, double*, int, int, void*) {
  cps::set_last_error("cps_modelfunc is a device-model token and is never executed on the host");
}
extern "C" void cps_jacmodelfunc(double*, double*, int, int, void*) {
  cps::set_last_error("cps_jacmodelfunc is a device-model token and is never executed on the host");
}
extern "C" void cps_modelfunc_image8(double*, double*, int, int, void*) {
  cps::set_last_error("cps_modelfunc_image8 is a device-model token and is never executed on the host");
}
extern "C" void cps_jacmodelfunc_image8(double*, double*, int, int, void*) {
  cps::set_last_error("cps_jacmodelfunc_image8 is a device-model token and is never executed on the host");
}

// registered host callbacks: pointers the same-signature entry points accept as names of a device model
namespace {
std::mutex g_reg_mutex;
std::vector<std::pair<void*, int>> g_reg_func, g_reg_jac;
int registered_model(const std::vector<std::pair<void*, int>>& tab, void* fn) {
  std::lock_guard<std::mutex> lock(g_reg_mutex);
  for (const auto& e : tab)
    if (e.first == fn) return e.second;
  return 0;
}
}  // namespace

extern "C" int cps_register_model(void (*func)(double*, double*, int, int, void*), int model_id) {
  if (!func || (model_id != CPS_MODEL_STEREO19 && model_id != CPS_MODEL_IMAGE8)) return CPS_ERR_INVALID;
  std::lock_guard<std::mutex> lock(g_reg_mutex);
  for (auto& e : g_reg_func)
    if (e.first == (void*)func) { e.second = model_id; return CPS_OK; }
  g_reg_func.emplace_back((void*)func, model_id);
  return CPS_OK;
}

extern "C" int cps_register_jacobian(void (*jacf)(double*, double*, int, int, void*), int model_id) {
  if (!jacf || (model_id != CPS_MODEL_STEREO19 && model_id != CPS_MODEL_IMAGE8)) return CPS_ERR_INVALID;
  std::lock_guard<std::mutex> lock(g_reg_mutex);
  for (auto& e : g_reg_jac)
    if (e.first == (void*)jacf) { e.second = model_id; return CPS_OK; }
  g_reg_jac.emplace_back((void*)jacf, model_id);
  return CPS_OK;
}

namespace {
// The levmar-named entry points carry no context argument: every calling thread gets its own, released when the
// thread ends (the main thread's goes before the static destructors, i.e. while the CUDA runtime is still up).
struct ShimCtxHolder {
  cps_lm_ctx* ctx = nullptr;
  ~ShimCtxHolder() {
    if (ctx) cps_lm_destroy(ctx);
  }
};
cps_lm_ctx* shim_ctx() {
  static thread_local ShimCtxHolder holder;
  if (!holder.ctx) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cps_lm_create(&holder.ctx, dev) != CPS_OK) holder.ctx = nullptr;
  }
  return holder.ctx;
}

int shim_fit(int variant, void (*func)(double*, double*, int, int, void*),
             void (*jacf)(double*, double*, int, int, void*), double* p, double* x, int m, int n, int itmax,
             double* opts, double* info, double* work, double* covar, void* adata) {
  int model;
  const int reg = registered_model(g_reg_func, (void*)func);
  if ((func == cps_modelfunc || reg == CPS_MODEL_STEREO19) && m == 19) model = CPS_MODEL_STEREO19;
  else if ((func == cps_modelfunc_image8 || reg == CPS_MODEL_IMAGE8) && m == 8) model = CPS_MODEL_IMAGE8;
  else {
    cps::set_last_error("unknown model callback: the device path has no CPU fallback for host callbacks");
    return CPS_ERR_UNSUPPORTED;
  }
  if (variant == CPS_LM_DER) {
    const int jreg = registered_model(g_reg_jac, (void*)jacf);
    if ((model == CPS_MODEL_STEREO19 && jacf != cps_jacmodelfunc && jreg != CPS_MODEL_STEREO19) ||
        (model == CPS_MODEL_IMAGE8 && jacf != cps_jacmodelfunc_image8 && jreg != CPS_MODEL_IMAGE8)) {
      cps::set_last_error("unknown Jacobian callback");
      return CPS_ERR_UNSUPPORTED;
    }
  }
  if (work || covar) {
    cps::set_last_error("work/covar must be NULL (the reference passes NULL, curveFittingClass.cpp:681)");
    return CPS_ERR_UNSUPPORTED;
  }
  if (!p || !x || !adata || n < 0) return CPS_ERR_INVALID;
  cps_lm_ctx* c = shim_ctx();
  if (!c) return CPS_ERR_CUDA;
  const cps_mydata* d = (const cps_mydata*)adata;
  int counts[4] = {d->num_left1, d->num_left2, d->num_right1, d->num_right2};  // measurement order, :99-117
  if (model == CPS_MODEL_IMAGE8) counts[0] = n / 2;
  long long off[2] = {0, n};
  int ret = 0;
  double linfo[CPS_LM_INFO_SZ];
  int rc = fit_batched_host(c, model, variant, 1, p, x, -1, d->t, off, counts, itmax, opts, linfo, &ret, nullptr);
  if (rc != CPS_OK) return rc;
  if (info) std::memcpy(info, linfo, sizeof linfo);
  return ret;
}
}  // namespace

extern "C" int cps_dlevmar_der(void (*func)(double*, double*, int, int, void*),
                               void (*jacf)(double*, double*, int, int, void*), double* p, double* x, int m, int n,
                               int itmax, double* opts, double* info, double* work, double* covar, void* adata) {
  return shim_fit(CPS_LM_DER, func, jacf, p, x, m, n, itmax, opts, info, work, covar, adata);
}

extern "C" int cps_dlevmar_dif(void (*func)(double*, double*, int, int, void*), double* p, double* x, int m, int n,
                               int itmax, double* opts, double* info, double* work, double* covar, void* adata) {
  return shim_fit(CPS_LM_DIF, func, nullptr, p, x, m, n, itmax, opts, info, work, covar, adata);
}
