// hadi — host layer (C++) behind the C ABI of include/hadi.h.
//
// Builds, for a batch of options, everything the reference's callers build before they launch
// compute_base_prices* / compute_jacobian* (grids, payoffs, per-maturity step tables —
// src/heston_calibration.cpp:2563-2660), packs it into flat device pools, launches the fused
// sm_100a kernel (hadi_kernel.cu) and runs the Levenberg-Marquardt driver
// (src/heston_calibration.cpp:2692-2831, src/jacobian_computation.cpp:20-195).
//
// sinh/asinh/exp are evaluated here, on the host, with libm: the oracle (the reference compiled for
// CPU) uses the same libm, and bit-equal grids are a precondition for bit-equal prices.
// There is no CPU solver in this file: every PDE solve goes through the CUDA kernel.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: the library is loaded with dlopen when a communicator is attached

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <new>
#include <vector>

#include "../../include/hadi.h"
#include "hadi_launch.h"

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  bool in_use = false;
  bool pinned_host = false;
};

}  // namespace

struct hadi_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 0;
  std::string err;
  long long launches = 0;
  long long exact_reruns = 0;
  long long h2d_bytes = 0, d2h_bytes = 0;  // cumulative transfer volume (bench.py reports it per step)
  std::vector<DevBuf> pool;  // caching allocator: device and pinned-host blocks are reused across batches
  // s-grids depend only on (m1, K, S0): cache them across calls (an LM run re-prices the same
  // strikes dozens of times; the reference likewise builds its GridViews once, before the loop).
  std::map<std::tuple<int, uint64_t, uint64_t>, std::shared_ptr<std::vector<double>>> s_cache;
  std::map<int, std::shared_ptr<std::vector<double>>> v_base;  // d*sinh(j*d_eta) per m2
  // Device-resident s-grids: a strike's grid is uploaded once and stays in HBM for the life of the context
  // (an option chain is re-priced many times with the same strikes: every LM iteration, every bench step), so
  // a steady-state call moves only the item descriptors.  One allocation that never moves: prepared batches keep
  // pointing into it.  When it is full, batches carry their grids with them as before.
  struct DevGrid { int off, idx_s; };
  double* d_spool = nullptr;
  size_t spool_cap = 0, spool_used = 0;   // doubles
  bool spool_tried = false;
  std::map<std::tuple<int, uint64_t, uint64_t>, DevGrid> s_dev;
  // In-library exchange (multi-GPU): an NCCL communicator bound to this context's device; all-gathers run on
  // `stream` behind the kernel whose epilogue wrote this rank's values straight into the gather buffer.
  ncclComm_t nccl = nullptr;
  int nccl_world = 1, nccl_rank = 0;
  double* d_gather = nullptr;   // [world][gather_cap] doubles
  double* h_gather = nullptr;   // pinned mirror
  size_t gather_cap = 0;        // doubles per rank
  // kernel plans per (m1, m2, scheme, few items, forced variant): the occupancy queries cost tens of microseconds
  std::map<std::tuple<int, int, int, int, std::string>, std::pair<int, HadiPlan>> plans;
};

struct hadi_batch {
  hadi_ctx* ctx = nullptr;
  int n_items = 0;
  int stride = 1;  // values per item (3 in HADI_MODE_JACOBIAN_INTERP)
  int n_hand = 0;  // hand-off slots of the split schedule (0: CTAs pull whole items from the counter)
  int m1 = 0, m2 = 0;
  HadiLaunch L{};
  HadiPlan plan{};
  int grid_ctas = 0;
  std::vector<int> bufs;  // indices into ctx->pool owned by this batch
  double* h_values = nullptr;  // pinned; one extra word past the values carries the re-solve counter
  long long reruns = 0;        // items of the last fetched launch that were re-solved with IEEE divisions
  // what hadi_batch_update_model needs: the descriptors as uploaded (sorted), the staging block and its layout
  std::vector<HadiItem> h_items;
  int mode = 0, item_begin = 0;
  double eps5[5] = {0, 0, 0, 0, 0};
  hadi_model model0{};
  char* h_stage = nullptr;
  char* d_stage = nullptr;
  size_t o_items = 0, o_v = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool launched = false;
};

namespace {

int fail(hadi_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}
int cuda_fail(hadi_ctx* ctx, cudaError_t e, const char* what) {
  return fail(ctx, HADI_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// ---- NCCL, loaded on demand ---------------------------------------------------------------------
// libhadi.so does not link NCCL: a single-GPU user never needs it, and in a PyTorch process the library torch
// already loaded (same soname) is the one dlopen returns, so both sides share one NCCL.
struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool ok() const { return lib && GetUniqueId && CommInitRank && CommDestroy && AllGather && GetErrorString; }
};
NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {getenv("HADI_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
      api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    }
  }
  return api;
}
int nccl_fail(hadi_ctx* ctx, ncclResult_t r, const char* what) {
  return fail(ctx, HADI_ERR_COMM, std::string(what) + ": " + (nccl_api().ok() ? nccl_api().GetErrorString(r) : "NCCL not loaded"));
}

uint64_t bits(double x) {
  uint64_t u;
  std::memcpy(&u, &x, sizeof u);
  return u;
}

int pool_get(hadi_ctx* ctx, size_t bytes, bool pinned_host) {
  bytes = (bytes + 255) & ~size_t(255);
  if (bytes == 0) bytes = 256;
  int best = -1;
  for (size_t k = 0; k < ctx->pool.size(); ++k) {
    DevBuf& b = ctx->pool[k];
    if (!b.in_use && b.pinned_host == pinned_host && b.bytes >= bytes && b.bytes <= 4 * bytes + (1 << 16))
      if (best < 0 || b.bytes < ctx->pool[best].bytes) best = (int)k;
  }
  if (best >= 0) {
    ctx->pool[best].in_use = true;
    return best;
  }
  DevBuf nb;
  nb.bytes = bytes;
  nb.pinned_host = pinned_host;
  cudaError_t e = pinned_host ? cudaMallocHost(&nb.p, bytes) : cudaMalloc(&nb.p, bytes);
  if (e != cudaSuccess) {
    ctx->err = std::string("allocation failed: ") + cudaGetErrorString(e);
    return -1;
  }
  nb.in_use = true;
  ctx->pool.push_back(nb);
  return (int)ctx->pool.size() - 1;
}

// ---- grids (src/grid.cpp:16-96; callers' constants S = 8K, c = K/5, V = 5, d = V/500) ----------
std::shared_ptr<std::vector<double>> s_grid(hadi_ctx* ctx, int m1, double K, double S0) {
  auto key = std::make_tuple(m1, bits(K), bits(S0));
  if (ctx) {
    auto it = ctx->s_cache.find(key);
    if (it != ctx->s_cache.end()) return it->second;
  }
  const double S = 8 * K, c = K / 5;
  auto g = std::make_shared<std::vector<double>>(m1 + 1);
  std::vector<double>& s = *g;
  const double lo = std::asinh(-K / c);
  const double dxi = (1.0 / m1) * (std::asinh((S - K) / c) - std::asinh(-K / c));
  for (int i = 0; i <= m1; ++i) {
    const double xi = lo + i * dxi;
    s[i] = K + c * std::sinh(xi);
  }
  s.push_back(S0);
  std::sort(s.begin(), s.end());
  s.pop_back();  // the reference drops the largest node (quirk Q1)
  if (ctx) {
    if (ctx->s_cache.size() > 200000) ctx->s_cache.clear();
    ctx->s_cache.emplace(key, g);
  }
  return g;
}

// src/grid_pod.hpp:25-73 with V = 5.0, d = 5.0/500 (hard-coded at every reference call site)
void v_grid(hadi_ctx* ctx, int m2, double V0, double* v) {
  const double V = 5.0, d = 5.0 / 500;
  std::shared_ptr<std::vector<double>> base;
  if (ctx) {
    auto it = ctx->v_base.find(m2);
    if (it != ctx->v_base.end()) base = it->second;
  }
  if (!base) {
    base = std::make_shared<std::vector<double>>(m2 + 1);
    const double deta = (1.0 / m2) * std::asinh(V / d);
    for (int j = 0; j <= m2; ++j) {
      const double xi = j * deta;
      (*base)[j] = d * std::sinh(xi);
    }
    if (ctx) ctx->v_base.emplace(m2, base);
  }
  std::vector<double> tmp(*base);
  tmp.push_back(V0);
  std::sort(tmp.begin(), tmp.end());
  for (int j = 0; j <= m2; ++j) v[j] = tmp[j];
}

int find_node(const double* x, int n, double x0) {
  for (int i = 0; i < n; ++i)
    if (std::fabs(x[i] - x0) < 1e-10) return i;
  return -1;
}

struct Geometry {
  int m1, m2, ld, n1, n2, pj;
};
Geometry geometry(int m1, int m2) {
  Geometry g;
  g.m1 = m1;
  g.m2 = m2;
  g.ld = (m1 + 1) | 1;  // odd pitch: conflict-free row AND column sweeps for 8-byte words
  g.n1 = (m1 + 1 + 3) & ~3;
  g.n2 = (m2 + 1 + 3) & ~3;
  g.pj = (m2 + 1 + 3) & ~3;
  return g;
}

bool valid_numerics(const hadi_numerics* num) {
  if (!num) return false;
  if (num->m1 < 4 || num->m2 < 4 || num->m1 > 4096 || num->m2 > 4096) return false;
  if (num->m2 > num->m1) return false;  // b1 lands on node (j, m1-j) only while m2 <= m1 (every reference caller: m1 = 2*m2)
  if (num->style != HADI_EUROPEAN && num->style != HADI_AMERICAN) return false;
  if (num->payoff != HADI_CALL && num->payoff != HADI_PUT) return false;
  if (num->num_dividends < 0) return false;
  if (num->num_dividends > 0 &&
      (!num->dividend_dates || !num->dividend_amounts || !num->dividend_percentages))
    return false;
  // opt-in extensions (parity unpinned): the reference defines Craig-Sneyd with its call boundary vectors only
  if (num->boundary != HADI_BC_REFERENCE_CALL && num->boundary != HADI_BC_PUT) return false;
  if (num->boundary == HADI_BC_PUT && num->scheme != HADI_DOUGLAS) return false;
  if (num->dividend_schedule != HADI_DIVIDENDS_DEVICE && num->dividend_schedule != HADI_DIVIDENDS_ALL) return false;
  return true;
}

// Split schedule (McNaughton's wrap-around rule for preemptive scheduling on identical machines): the batch is
// n solves of N_k time steps on `slots` persistent CTAs.  Whole solves per CTA quantise the makespan to a whole
// number of solves per CTA (500 solves on 296 CTAs take as long as 592); cutting the one solve that straddles
// the end of a CTA's share in two makes every CTA finish after T = total / slots steps.  The LAST steps of a
// cut solve close the list of CTA b, its FIRST steps open the list of CTA b + 1, so the state (U, lambda)
// is ready long before it is needed whenever T >= N + 2 set-ups.  Costs are in time steps; `setup` is what a
// segment pays before its first step (tables, factorisation; measured ~1.3 steps at 101x51).
struct SplitSchedule {
  std::vector<HadiSegment> segs;
  std::vector<int> off;   // [slots + 1]
  int n_hand = 0;
};
// One pass of the rule for a given target T; returns the load (steps) of the heaviest CTA.
double fill_split_schedule(const std::vector<HadiItem>& items, int slots, double setup, double T, SplitSchedule* out) {
  const int n = (int)items.size();
  const int min_seg = 3;   // a cut that leaves fewer steps than this on either side is not worth a set-up
  out->segs.clear();
  out->off.assign(1, 0);
  out->n_hand = 0;
  int slot = 0;
  double used = 0.0, heaviest = 0.0;
  auto close_slot = [&]() {
    heaviest = std::max(heaviest, used);
    out->off.push_back((int)out->segs.size());
    ++slot;
    used = 0.0;
  };
  for (int k = 0; k < n; ++k) {
    const int N = items[k].N;
    const double room = T - used - setup;
    if (slot >= slots - 1 || room >= N) {
      out->segs.push_back(HadiSegment{k, 1, N, -1, -1, 0, 0, 0});
      used += setup + N;
      continue;
    }
    int tail = (int)room;   // steps of this solve that still fit
    if (tail > N - min_seg) tail = N - min_seg;   // nearly all of it fits: leave min_seg steps for the next CTA
    if (tail < min_seg) {
      close_slot();
      out->segs.push_back(HadiSegment{k, 1, N, -1, -1, 0, 0, 0});
      used += setup + N;
      continue;
    }
    const int h = out->n_hand++;
    out->segs.push_back(HadiSegment{k, N - tail + 1, N, h, -1, 0, 0, 0});   // closes this CTA's list
    used += setup + tail;
    close_slot();
    out->segs.push_back(HadiSegment{k, 1, N - tail, -1, h, 0, 0, 0});       // opens the next CTA's list
    used += setup + (N - tail);
  }
  heaviest = std::max(heaviest, used);
  while ((int)out->off.size() < slots + 1) out->off.push_back((int)out->segs.size());
  return heaviest;
}
bool build_split_schedule(const std::vector<HadiItem>& items, int slots, double setup, SplitSchedule* out) {
  const int n = (int)items.size();
  if (slots < 2 || n <= slots) return false;
  double total = 0.0, longest = 0.0;
  for (const HadiItem& it : items) {
    total += it.N + setup;
    longest = std::max(longest, it.N + setup);
  }
  // every CTA but the last may end with a cut: one more set-up each.  Cuts fall on whole steps and slivers
  // shorter than min_seg are left unused, so the last CTA collects what the others could not place: raise T
  // until it is no heavier than the rest.
  double T = std::max(longest, (total + (slots - 1) * setup) / slots);
  for (int iter = 0; iter < 64; ++iter) {
    const double heaviest = fill_split_schedule(items, slots, setup, T, out);
    if (heaviest <= T + 1e-9) break;
    T += std::max(0.125, (heaviest - T) / slots);
  }
  return out->n_hand > 0;
}

// work items per option: base + one per bumped parameter
int n_columns(int mode) {
  switch (mode) {
    case HADI_MODE_JACOBIAN: return 6;           // base, kappa, eta, sigma, rho, v0        (forward differences)
    case HADI_MODE_JACOBIAN_INTERP: return 5;    // base, kappa, eta, sigma, rho            (v0 by interpolation)
    case HADI_MODE_JACOBIAN_CENTRAL: return 11;  // base, +5 bumps, -5 bumps                (central differences)
    default: return 1;
  }
}
bool valid_mode(int mode) {
  return mode == HADI_MODE_PRICE || mode == HADI_MODE_JACOBIAN || mode == HADI_MODE_JACOBIAN_INTERP ||
         mode == HADI_MODE_JACOBIAN_CENTRAL;
}
// The v-rows bracketing V0 + eps on the base v-grid and the interpolation weight
// (src/device_solver.cpp:1735-1754: first i with v_i <= V0+eps <= v_{i+1}; all zero when there is none).
void v0_bracket(const double* v, int m2, double v0_pert, int* lo, int* hi, double* weight) {
  *lo = 0; *hi = 0; *weight = 0.0;
  for (int i = 0; i < m2; ++i) {
    if (v[i] <= v0_pert && v0_pert <= v[i + 1]) {
      *lo = i;
      *hi = i + 1;
      *weight = (v0_pert - v[i]) / (v[i + 1] - v[i]);
      break;
    }
  }
}

}  // namespace

// "nothing throws across the ABI" (include/hadi.h): the entry points that size std::vector / std::map from caller input are
// function-try-blocks
#define HADI_CATCH \
  catch (const std::bad_alloc&) { return HADI_ERR_NOMEM; } \
  catch (...) { return HADI_ERR_ARG; }

extern "C" {

const char* hadi_version(void) { return "hadi 0.1 (sm_100a)"; }

int hadi_create(hadi_ctx** out, int device) try {
  if (!out) return HADI_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return HADI_ERR_CUDA;
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return HADI_ERR_CUDA;
  hadi_ctx* ctx = new hadi_ctx();
  ctx->device = device;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return HADI_ERR_CUDA;
  }
  if (cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return HADI_ERR_CUDA;
  }
  *out = ctx;
  return HADI_OK;
} HADI_CATCH

void hadi_destroy(hadi_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (DevBuf& b : ctx->pool) {
    if (b.pinned_host)
      cudaFreeHost(b.p);
    else
      cudaFree(b.p);
  }
  if (ctx->d_spool) cudaFree(ctx->d_spool);
  if (ctx->nccl && nccl_api().ok()) nccl_api().CommDestroy(ctx->nccl);
  if (ctx->d_gather) cudaFree(ctx->d_gather);
  if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

// ---- in-library multi-GPU exchange -------------------------------------------------
int hadi_nccl_unique_id(void* id128) try {
  if (!id128) return HADI_ERR_ARG;
  if (!nccl_api().ok()) return HADI_ERR_COMM;
  ncclUniqueId id;
  if (nccl_api().GetUniqueId(&id) != ncclSuccess) return HADI_ERR_COMM;
  static_assert(sizeof(id) == HADI_NCCL_ID_BYTES, "ncclUniqueId is 128 bytes");
  std::memcpy(id128, &id, sizeof id);
  return HADI_OK;
} HADI_CATCH

int hadi_comm_init(hadi_ctx* ctx, int world, int rank, const void* id128) try {
  if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return HADI_ERR_ARG;
  if (ctx->nccl) return fail(ctx, HADI_ERR_ARG, "a communicator is already attached");
  if (!nccl_api().ok()) return fail(ctx, HADI_ERR_COMM, "libnccl.so.2 could not be loaded (set HADI_NCCL_LIB)");
  cudaSetDevice(ctx->device);
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof id);
  const ncclResult_t r = nccl_api().CommInitRank(&ctx->nccl, world, id, rank);
  if (r != ncclSuccess) {
    ctx->nccl = nullptr;
    return nccl_fail(ctx, r, "ncclCommInitRank");
  }
  ctx->nccl_world = world;
  ctx->nccl_rank = rank;
  return HADI_OK;
} HADI_CATCH

void hadi_comm_finalize(hadi_ctx* ctx) {
  if (!ctx || !ctx->nccl) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (nccl_api().ok()) nccl_api().CommDestroy(ctx->nccl);
  ctx->nccl = nullptr;
  ctx->nccl_world = 1;
  ctx->nccl_rank = 0;
}

int hadi_comm_world(const hadi_ctx* ctx) { return ctx ? ctx->nccl_world : 0; }
int hadi_comm_rank(const hadi_ctx* ctx) { return ctx ? ctx->nccl_rank : -1; }

const char* hadi_last_error(const hadi_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }
long long hadi_kernel_launches(const hadi_ctx* ctx) { return ctx ? ctx->launches : 0; }
long long hadi_exact_reruns(const hadi_ctx* ctx) { return ctx ? ctx->exact_reruns : 0; }
int hadi_transfer_bytes(const hadi_ctx* ctx, long long* h2d, long long* d2h) try {
  if (!ctx) return HADI_ERR_ARG;
  if (h2d) *h2d = ctx->h2d_bytes;
  if (d2h) *d2h = ctx->d2h_bytes;
  return HADI_OK;
} HADI_CATCH

int hadi_grid(int m1, int m2, double K, double S0, double V0, double* s, double* v) try {
  if (m1 < 1 || m2 < 1 || !s || !v) return HADI_ERR_ARG;
  auto g = s_grid(nullptr, m1, K, S0);
  std::copy(g->begin(), g->end(), s);
  v_grid(nullptr, m2, V0, v);
  return HADI_OK;
} HADI_CATCH

double hadi_bs_call(double S, double K, double r, double vol, double T) {
  // src/bs.hpp:44-55
  const double sqrt_T = std::sqrt(T);
  const double log_SK = std::log(S / K);
  const double vol_sqrt_T = vol * sqrt_T;
  const double d1 = (log_SK + (r + 0.5 * vol * vol) * T) / vol_sqrt_T;
  const double d2 = d1 - vol_sqrt_T;
  return S * std::erfc(-d1 / std::sqrt(2.0)) / 2.0 - K * std::exp(-r * T) * std::erfc(-d2 / std::sqrt(2.0)) / 2.0;
}

int hadi_item_costs(const hadi_numerics* num, int n, const hadi_point* points, int mode, int* costs) try {
  if (!valid_numerics(num) || n < 0 || (n > 0 && (!points || !costs))) return HADI_ERR_ARG;
  const int nc = n_columns(mode);
  const int P = (num->m1 + 1) * (num->m2 + 1);
  for (int k = 0; k < n; ++k)
    for (int c = 0; c < nc; ++c) costs[k * nc + c] = points[k].time_steps * P;
  return HADI_OK;
} HADI_CATCH

int hadi_partition(int n_items, const int* costs, int world, int rank, int* begin, int* end) try {
  if (n_items < 0 || world <= 0 || rank < 0 || rank >= world || !begin || !end) return HADI_ERR_ARG;
  // contiguous blocks; boundary r is the first item whose cost prefix reaches r/world of the total
  long long total = 0;
  for (int k = 0; k < n_items; ++k) total += costs ? costs[k] : 1;
  auto boundary = [&](int r) {
    if (r <= 0) return 0;
    if (r >= world) return n_items;
    const long long target = (total * r + world - 1) / world;
    long long acc = 0;
    int k = 0;
    while (k < n_items && acc < target) {
      acc += costs ? costs[k] : 1;
      ++k;
    }
    return k;
  };
  *begin = boundary(rank);
  *end = boundary(rank + 1);
  return HADI_OK;
} HADI_CATCH

// Inspection / test aid: the split schedule hadi_batch_create builds for n solves of time_steps[k] steps
// (in the order given) on `slots` persistent CTAs.
int hadi_plan_schedule(int n, const int* time_steps, int slots, double setup, int max_segments, int* seg5,
                       int* slot_off, double* heaviest_steps) try {
  if (n < 0 || (n > 0 && !time_steps) || slots < 1 || !seg5 || !slot_off) return HADI_ERR_ARG;
  std::vector<HadiItem> items((size_t)n);
  for (int k = 0; k < n; ++k) {
    if (time_steps[k] < 1) return HADI_ERR_ARG;
    std::memset(&items[k], 0, sizeof(HadiItem));
    items[k].N = time_steps[k];
  }
  SplitSchedule sc;
  if (!build_split_schedule(items, slots, setup, &sc)) return 0;
  if ((int)sc.segs.size() > max_segments) return HADI_ERR_ARG;
  for (size_t q = 0; q < sc.segs.size(); ++q) {
    seg5[5 * q + 0] = sc.segs[q].item; seg5[5 * q + 1] = sc.segs[q].n0; seg5[5 * q + 2] = sc.segs[q].n1;
    seg5[5 * q + 3] = sc.segs[q].hin;  seg5[5 * q + 4] = sc.segs[q].hout;
  }
  for (int b = 0; b <= slots; ++b) slot_off[b] = sc.off[b];
  if (heaviest_steps) {
    double hv = 0.0;
    for (int b = 0; b < slots; ++b) {
      double u = 0.0;
      for (int q = sc.off[b]; q < sc.off[b + 1]; ++q) u += setup + (sc.segs[q].n1 - sc.segs[q].n0 + 1);
      hv = std::max(hv, u);
    }
    *heaviest_steps = hv;
  }
  return (int)sc.segs.size();
} HADI_CATCH

int hadi_jacobian_assemble(int n, const double* v, double eps, double* J, double* base) try {
  if (n < 0 || !v || !J || !base) return HADI_ERR_ARG;
  for (int k = 0; k < n; ++k) {
    const double b = v[6 * k];
    base[k] = b;
    for (int c = 0; c < 5; ++c) J[5 * k + c] = (v[6 * k + 1 + c] - b) / eps;
  }
  return HADI_OK;
} HADI_CATCH

int hadi_jacobian_v0_weight(int m2, double V0, double eps_v0, int* lower, int* upper, double* weight) try {
  if (m2 < 1 || !lower || !upper || !weight) return HADI_ERR_ARG;
  std::vector<double> v((size_t)m2 + 1);
  v_grid(nullptr, m2, V0, v.data());
  v0_bracket(v.data(), m2, V0 + eps_v0, lower, upper, weight);
  return HADI_OK;
} HADI_CATCH

// Jacobian rows from the item values of any Jacobian mode (layouts: include/hadi.h).
int hadi_jacobian_assemble_ex(int n, int mode, const double* v, const double* eps5, double v0_weight, double* J,
                              double* base) try {
  if (n < 0 || !v || !eps5 || !J || !base) return HADI_ERR_ARG;
  if (mode == HADI_MODE_JACOBIAN) {
    // src/jacobian_computation.cpp:330,361
    for (int k = 0; k < n; ++k) {
      const double b = v[6 * k];
      base[k] = b;
      for (int c = 0; c < 5; ++c) J[5 * k + c] = (v[6 * k + 1 + c] - b) / eps5[c];
    }
  } else if (mode == HADI_MODE_JACOBIAN_INTERP) {
    // src/device_solver.cpp:1806-1818: the V0 column from the base solve, interpolated linearly in v
    for (int k = 0; k < n; ++k) {
      const double* o = v + (size_t)15 * k;   // 5 items x {price, U(S0, v_lower), U(S0, v_upper)}
      const double b = o[0];
      base[k] = b;
      for (int c = 0; c < 4; ++c) J[5 * k + c] = (o[3 * (1 + c)] - b) / eps5[c];
      const double price_lower = o[1], price_upper = o[2];
      const double pert_price = price_lower + v0_weight * (price_upper - price_lower);
      J[5 * k + 4] = (pert_price - b) / eps5[4];
    }
  } else if (mode == HADI_MODE_JACOBIAN_CENTRAL) {
    for (int k = 0; k < n; ++k) {
      const double* o = v + (size_t)11 * k;
      base[k] = o[0];
      for (int c = 0; c < 5; ++c) J[5 * k + c] = (o[1 + c] - o[6 + c]) / (2.0 * eps5[c]);
    }
  } else {
    return HADI_ERR_ARG;
  }
  return HADI_OK;
} HADI_CATCH

// ------------------------------------------------------------------------------------------------
int hadi_batch_create(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                      const hadi_point* points, int mode, double eps, int item_begin, int item_end,
                      hadi_batch** out) {
  const double eps5[5] = {eps, eps, eps, eps, eps};
  return hadi_batch_create_ex(ctx, model, num, n, points, mode, eps5, item_begin, item_end, out);
}

int hadi_batch_create_ex(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                         const hadi_point* points, int mode, const double* eps5, int item_begin, int item_end,
                         hadi_batch** out) try {
  if (!ctx || !out) return HADI_ERR_ARG;
  *out = nullptr;
  if (!model || !valid_numerics(num) || n < 0 || (n > 0 && !points) || !eps5) return fail(ctx, HADI_ERR_ARG, "bad argument");
  if (!valid_mode(mode)) return fail(ctx, HADI_ERR_ARG, "bad mode");
  if (mode == HADI_MODE_JACOBIAN_CENTRAL && !(model->V0 - eps5[4] > 0.0))
    return fail(ctx, HADI_ERR_ARG, "central differences need V0 - eps > 0");
  if (num->scheme != HADI_DOUGLAS && num->scheme != HADI_CRAIG_SNEYD && num->scheme != HADI_MODIFIED_CRAIG_SNEYD &&
      num->scheme != HADI_HUNDSDORFER_VERWER)
    return fail(ctx, HADI_ERR_ARG, "unknown scheme");
  // the reference defines Craig-Sneyd for European options without dividends only (src/solver.hpp:781)
  if (num->scheme != HADI_DOUGLAS && (num->style != HADI_EUROPEAN || num->num_dividends > 0))
    return fail(ctx, HADI_ERR_ARG, "Craig-Sneyd: European options without dividends only");
  const int nc = n_columns(mode);
  const int total_items = n * nc;
  if (item_end < 0) item_end = total_items;
  if (item_begin < 0 || item_begin > item_end || item_end > total_items) return fail(ctx, HADI_ERR_ARG, "bad item range");
  cudaSetDevice(ctx->device);

  const Geometry g = geometry(num->m1, num->m2);
  const int m1 = g.m1, m2 = g.m2;
  const int P = (m1 + 1) * (m2 + 1);
  HadiPlan plan;
  const char* forced_variant = getenv("HADI_FORCE_VARIANT");
  const int n_it_plan = (item_end < 0 ? n * n_columns(mode) : item_end) - item_begin;
  const bool many_items = n_it_plan >= 6 * 148;   // fills the slots of the small-CTA instantiation (51 x 26: variant 11)
  const auto plan_key = std::make_tuple(num->m1, num->m2, num->scheme,
                                        (int)(n_it_plan * HADI_CLUSTER <= 148 && num->num_dividends == 0) + 2 * (int)many_items,
                                        std::string(forced_variant ? forced_variant : "") + "|" +
                                            std::string(getenv("HADI_NO_DUO") ? getenv("HADI_NO_DUO") : ""));
  const auto plan_hit = ctx->plans.find(plan_key);
  const bool forced_wide = forced_variant && atoi(forced_variant) == HADI_WIDE_VARIANT;
  bool no_base_plan = false;
  if (forced_variant && atoi(forced_variant) == 7 && num->num_dividends > 0)   // the planner never pairs them; a forced run must not skip the jumps silently
    return fail(ctx, HADI_ERR_ARG, "the cluster kernel (variant 7) does not take dividend jumps");
  if (forced_wide) {
    plan.global_state = true;   // filled in below
  } else if (plan_hit != ctx->plans.end() && plan_hit->second.first == 0) {
    plan = plan_hit->second.second;
  } else {
    const bool cs = num->scheme != HADI_DOUGLAS;   // the Craig-Sneyd family runs on the global-state kernels
    // Few large solves run on the wide kernel (below).  The thread-block-cluster kernel of round 1 is no longer a
    // planner choice — only HADI_FORCE_VARIANT=7 reaches it (hadi_douglas_plan honours the forced id itself); where the
    // wide kernel does not take a grid, few solves run one CTA per solve like many.
    const int n_it = (item_end < 0 ? n * n_columns(mode) : item_end) - item_begin;
    (void)n_it;
    const bool few = false;
    int prc = hadi_douglas_plan(ctx->device, m1, m2, g.ld, g.n1, g.n2, g.pj, cs, &plan, few, many_items);
    // grids beyond shared memory run on the global-state kernel (working set in L2-resident scratch)
    if (prc < 0 && !cs) prc = hadi_douglas_plan(ctx->device, m1, m2, g.ld, g.n1, g.n2, g.pj, true, &plan, few);
    // a Douglas grid that only the global-state kernel takes: same choice between one CTA and one cluster per solve
    if (prc == 0 && plan.global_state && plan.cluster <= 1 && few) {
      HadiPlan cplan;
      if (hadi_douglas_plan(ctx->device, m1, m2, g.ld, g.n1, g.n2, g.pj, true, &cplan, true) == 0 && cplan.cluster > 1)
        plan = cplan;
    }
    if (prc > 0) return cuda_fail(ctx, (cudaError_t)prc, "kernel plan");
    if (prc < 0) no_base_plan = true;   // no one-CTA kernel takes this grid: the wide kernel may (it needs m2 + 1 <= 511 only)
    else ctx->plans[plan_key] = std::make_pair(0, plan);
  }
  // A few solves that need the global working set (large grids, the Craig-Sneyd family): the
  // wide kernel spreads each over a team of co-resident CTAs (hadi_wide.cu).  HADI_WIDE_MAX_ITEMS moves the threshold
  // (0 disables); HADI_FORCE_VARIANT=9 takes it for any grid and batch size.
  {
    const char* wm = getenv("HADI_WIDE_MAX_ITEMS");
    // measured on B200 (tools/time_wide.py): 401 x 201 Craig-Sneyd breaks even with the one-CTA kernel near 40 solves,
    // 101 x 51 near 80 (the smaller the grid, the more of its lines fit one CTA's shared memory at once)
    const int wide_max = wm ? atoi(wm) : (P > 16384 ? HADI_WIDE_MAX_ITEMS_DEFAULT : 2 * HADI_WIDE_MAX_ITEMS_DEFAULT);
    // Grids so small that the A2 assembly scratch (TS_COUNT rows of n2 doubles) does not fit the Y array the other
    // kernels borrow for it (roughly m1 < 20) also go to the wide kernel, whatever the batch: it has its own arena.
    const bool tiny_grid = (size_t)TS_COUNT * g.n2 > (size_t)(m2 + 1) * g.ld;
    if (n_it_plan >= 1 && (forced_wide || tiny_grid || (no_base_plan && !forced_variant) ||
                           (!forced_variant && !no_base_plan && plan.global_state && n_it_plan <= wide_max))) {
      const auto wkey = std::make_tuple(num->m1, num->m2, num->scheme, 2, std::string("wide"));
      auto wh = ctx->plans.find(wkey);
      if (wh == ctx->plans.end()) {
        HadiPlan wp;
        const int wrc = hadi_wide_plan(ctx->device, m1, m2, g.ld, g.n1, g.n2, g.pj, &wp);
        wh = ctx->plans.emplace(wkey, std::make_pair(wrc, wp)).first;
      }
      if (wh->second.first == 0) {
        plan = wh->second.second;
        plan.cluster = hadi_wide_team(n_it_plan, plan.sm_count, m1, m2);
      } else if (forced_wide || tiny_grid || no_base_plan) {
        return fail(ctx, HADI_ERR_SMEM, no_base_plan ? "grid too large: m1+1 <= 1024, m2+1 <= 511 and the coefficient tables must fit shared memory"
                                                     : "the wide kernel does not take this grid");
      }
      no_base_plan = false;
    }
    if (no_base_plan) return fail(ctx, HADI_ERR_SMEM, "grid too large: m1+1 <= 1024 and the coefficient tables must fit shared memory");
  }

  std::unique_ptr<hadi_batch> b(new hadi_batch());
  b->ctx = ctx;
  b->m1 = m1;
  b->m2 = m2;
  const int n_items = item_end - item_begin;
  b->n_items = n_items;

  // ---- host-side descriptors --------------------------------------------------------------
  std::vector<HadiItem> items(n_items);
  std::vector<double> s_pool, v_pool, e_pool;
  std::map<std::pair<uint64_t, uint64_t>, int> s_index;       // (K, S0) -> offset
  // device-resident grid pool: usable when the grids this batch adds still fit
  std::map<std::tuple<int, uint64_t, uint64_t>, hadi_ctx::DevGrid> pending;   // grids this call uploads
  bool use_dev_pool = false;
  if (!getenv("HADI_NO_GRID_CACHE")) {
    if (!ctx->spool_tried) {
      ctx->spool_tried = true;
      const size_t cap = (size_t)4 << 20;   // 4 Mi doubles = 32 MiB: ~41 000 grids of 101 nodes
      if (cudaMalloc((void**)&ctx->d_spool, cap * sizeof(double)) == cudaSuccess) ctx->spool_cap = cap;
      else { ctx->d_spool = nullptr; cudaGetLastError(); }
    }
    if (ctx->d_spool) {
      size_t need = 0;
      std::map<std::tuple<int, uint64_t, uint64_t>, char> seen;
      const int nc_ = n_columns(mode);
      const int k0 = item_begin / nc_, k1 = n_items > 0 ? (item_end - 1) / nc_ : k0 - 1;
      for (int k = k0; k <= k1; ++k) {
        const auto dkey = std::make_tuple(m1, bits(points[k].strike), bits(model->S0));
        if (ctx->s_dev.find(dkey) == ctx->s_dev.end() && seen.emplace(dkey, 1).second) need += (size_t)m1 + 1;
      }
      use_dev_pool = ctx->spool_used + need <= ctx->spool_cap;
    }
  }
  std::map<std::pair<uint64_t, int>, int> e_index;            // (dt, N) -> offset
  // three v-grids at most: V0, V0 + eps (src/jacobian_computation.cpp:339) and, for central differences, V0 - eps
  v_pool.resize((size_t)3 * (m2 + 1));
  v_grid(ctx, m2, model->V0, v_pool.data());
  const double V0p = model->V0 + eps5[4], V0m = model->V0 - eps5[4];
  v_grid(ctx, m2, V0p, v_pool.data() + (m2 + 1));
  v_grid(ctx, m2, mode == HADI_MODE_JACOBIAN_CENTRAL ? V0m : model->V0, v_pool.data() + 2 * (m2 + 1));
  int idx_v0 = find_node(v_pool.data(), m2 + 1, model->V0);
  if (idx_v0 < 0) idx_v0 = 0;  // find_v0_index returns 0 when nothing matches (src/grid_pod.hpp:76-87)
  int idx_v1 = find_node(v_pool.data() + (m2 + 1), m2 + 1, V0p);
  if (idx_v1 < 0) idx_v1 = 0;
  int idx_v2 = find_node(v_pool.data() + 2 * (m2 + 1), m2 + 1, V0m);
  if (idx_v2 < 0) idx_v2 = 0;
  int br_lo = 0, br_hi = 0;
  double br_w = 0.0;
  v0_bracket(v_pool.data(), m2, V0p, &br_lo, &br_hi, &br_w);
  (void)br_w;  // applied by hadi_jacobian_assemble_ex
  b->stride = (mode == HADI_MODE_JACOBIAN_INTERP) ? 3 : 1;

  for (int q = 0; q < n_items; ++q) {
    const int item = item_begin + q;
    const int k = item / nc, col = item % nc;
    const hadi_point& pt = points[k];
    if (pt.time_steps < 1 || !(pt.delta_t > 0)) return fail(ctx, HADI_ERR_ARG, "bad time stepping");
    HadiItem it;
    std::memset(&it, 0, sizeof it);
    it.kappa = model->kappa;
    it.eta = model->eta;
    it.sigma = model->sigma;
    it.rho = model->rho;
    // bumps: src/jacobian_computation.cpp:299-304 (param + eps), :339 (grid for V0 + eps); columns 6..10
    // are the downward bumps of the central-difference mode
    if (col == 1) it.kappa += eps5[0];
    if (col == 2) it.eta += eps5[1];
    if (col == 3) it.sigma += eps5[2];
    if (col == 4) it.rho += eps5[3];
    if (col == 6) it.kappa -= eps5[0];
    if (col == 7) it.eta -= eps5[1];
    if (col == 8) it.sigma -= eps5[2];
    if (col == 9) it.rho -= eps5[3];
    it.r_d = model->r_d;
    it.r_f = model->r_f;
    it.dt = pt.delta_t;
    it.theta = num->theta;
    it.K = pt.strike;
    it.N = pt.time_steps;
    it.ef = std::exp(-model->r_f * pt.delta_t * (pt.time_steps - 1));
    it.style = num->style;
    it.payoff = num->payoff;
    it.nd = num->num_dividends;
    it.bc = num->boundary;
    it.div_all = num->dividend_schedule;
    if (use_dev_pool) {
      const auto dkey = std::make_tuple(m1, bits(pt.strike), bits(model->S0));
      auto dit = ctx->s_dev.find(dkey);
      if (dit == ctx->s_dev.end()) {
        dit = pending.find(dkey);
        if (dit == pending.end()) {
          auto gs = s_grid(ctx, m1, pt.strike, model->S0);
          hadi_ctx::DevGrid dg;
          dg.off = (int)(ctx->spool_used + s_pool.size());
          dg.idx_s = find_node(gs->data(), m1 + 1, model->S0);
          if (dg.idx_s < 0) return fail(ctx, HADI_ERR_GRID, "S0 is not a node of the s-grid");
          s_pool.insert(s_pool.end(), gs->begin(), gs->end());
          dit = pending.emplace(dkey, dg).first;
        }
      }
      it.s_off = dit->second.off;
      it.idx_s = dit->second.idx_s;
    } else {
      auto skey = std::make_pair(bits(pt.strike), bits(model->S0));
      auto sit = s_index.find(skey);
      if (sit == s_index.end()) {
        auto gs = s_grid(ctx, m1, pt.strike, model->S0);
        const int off = (int)s_pool.size();
        s_pool.insert(s_pool.end(), gs->begin(), gs->end());
        s_index.emplace(skey, off);
        it.s_off = off;
      } else {
        it.s_off = sit->second;
      }
      it.idx_s = find_node(s_pool.data() + it.s_off, m1 + 1, model->S0);
      if (it.idx_s < 0) return fail(ctx, HADI_ERR_GRID, "S0 is not a node of the s-grid");
    }
    it.v_off = (col == 5) ? (m2 + 1) : (col == 10) ? 2 * (m2 + 1) : 0;
    it.idx_v = (col == 5) ? idx_v1 : (col == 10) ? idx_v2 : idx_v0;
    it.aux = br_lo | (br_hi << 16);
    auto ekey = std::make_pair(bits(pt.delta_t), pt.time_steps);
    auto eit = e_index.find(ekey);
    if (eit == e_index.end()) {
      const int off = (int)e_pool.size();
      for (int nn = 0; nn <= pt.time_steps; ++nn) e_pool.push_back(std::exp(model->r_f * pt.delta_t * nn));
      // discount table of the put-correct boundary set (HADI_BC_PUT): exp(-r_d*dt*n) behind the r_f table
      if (num->boundary == HADI_BC_PUT)
        for (int nn = 0; nn <= pt.time_steps; ++nn) e_pool.push_back(std::exp(-model->r_d * pt.delta_t * nn));
      e_index.emplace(ekey, off);
      it.e_off = off;
    } else {
      it.e_off = eit->second;
    }
    it.out = q;
    it.cost = pt.time_steps * P;
    items[q] = it;
  }
  // longest items first (the persistent CTAs pull items in order)
  std::stable_sort(items.begin(), items.end(), [](const HadiItem& a, const HadiItem& c) { return a.cost > c.cost; });

  // split schedule for the shared-memory resident kernels (the ring-fed and global-state kernels keep whole items)
  SplitSchedule sched;
  bool use_split = false;
  {
    const int slots = plan.ctas_per_sm * plan.sm_count;
    const char* ns = getenv("HADI_NO_SPLIT");
    const char* su = getenv("HADI_SPLIT_SETUP");
    const double setup = su ? atof(su) : 1.5;
    if (!(ns && atoi(ns) != 0) && !plan.global_state && plan.cluster <= 1 && (plan.variant <= 3 || plan.variant == 8 || plan.variant == 11) && !getenv("HADI_MAX_CTAS"))
      use_split = build_split_schedule(items, slots, setup, &sched);
  }

  // ---- device buffers ------------------------------------------------------------------------
  const int nd = num->num_dividends;
  const size_t bytes_items = sizeof(HadiItem) * (size_t)std::max(n_items, 1);
  const size_t bytes_s = sizeof(double) * std::max<size_t>(s_pool.size(), 1);
  const size_t bytes_v = sizeof(double) * v_pool.size();
  const size_t bytes_e = sizeof(double) * std::max<size_t>(e_pool.size(), 1);
  const size_t bytes_d = sizeof(double) * (size_t)std::max(3 * nd, 1);
  const size_t bytes_sg = use_split ? sizeof(HadiSegment) * sched.segs.size() + sizeof(int) * sched.off.size() : 0;
  const size_t staging = bytes_items + bytes_s + bytes_v + bytes_e + bytes_d + bytes_sg + 8 * 256;

  auto take = [&](size_t bytes, bool pinned) -> void* {
    const int id = pool_get(ctx, bytes, pinned);
    if (id < 0) return nullptr;
    b->bufs.push_back(id);
    return ctx->pool[id].p;
  };
  auto release_all = [&]() {
    for (int id : b->bufs) ctx->pool[id].in_use = false;
  };
  char* h_stage = (char*)take(staging, true);
  char* d_stage = (char*)take(staging, false);
  b->h_values = (double*)take(sizeof(double) * ((size_t)std::max(n_items, 1) * b->stride + 1), true);
  double* d_values = (double*)take(sizeof(double) * ((size_t)std::max(n_items, 1) * b->stride + 1), false);
  int* d_counter = (int*)take(sizeof(int) * HADI_COUNTER_INTS, false);

  b->plan = plan;
  b->grid_ctas = std::max(1, std::min(n_items, plan.ctas_per_sm * plan.sm_count));
  if (plan.cluster > 1) b->grid_ctas = plan.cluster * std::max(1, std::min(n_items, plan.sm_count / plan.cluster));
  if (const char* cap = getenv("HADI_MAX_CTAS"))  // development aid: cap the persistent grid
    if (atoi(cap) > 0) b->grid_ctas = std::min(b->grid_ctas, atoi(cap));
  const size_t stride = hadi_scratch_layout(m1, m2, g.ld, g.pj, plan.global_state, num->scheme != HADI_DOUGLAS).total;
  double* d_scratch = (double*)take(sizeof(double) * stride * (size_t)(b->grid_ctas / std::max(1, plan.cluster)), false);
  if (!h_stage || !d_stage || !b->h_values || !d_values || !d_counter || !d_scratch) {
    release_all();
    return HADI_ERR_NOMEM;
  }

  size_t off = 0;
  auto put = [&](const void* src, size_t bytes) -> size_t {
    const size_t at = off;
    if (bytes) std::memcpy(h_stage + at, src, bytes);
    off = (off + bytes + 255) & ~size_t(255);
    return at;
  };
  const size_t o_items = put(items.data(), sizeof(HadiItem) * (size_t)n_items);
  b->h_items = items;
  b->mode = mode;
  b->item_begin = item_begin;
  for (int c = 0; c < 5; ++c) b->eps5[c] = eps5[c];
  b->model0 = *model;
  b->h_stage = h_stage;
  b->d_stage = d_stage;
  b->o_items = o_items;
  const size_t o_s = put(s_pool.data(), sizeof(double) * s_pool.size());
  const size_t o_v = put(v_pool.data(), bytes_v);
  b->o_v = o_v;
  const size_t o_e = put(e_pool.data(), sizeof(double) * e_pool.size());
  std::vector<double> dv((size_t)std::max(3 * nd, 1), 0.0);
  for (int k = 0; k < nd; ++k) {
    dv[k] = num->dividend_dates[k];
    dv[nd + k] = num->dividend_amounts[k];
    dv[2 * nd + k] = num->dividend_percentages[k];
  }
  const size_t o_d = put(dv.data(), sizeof(double) * dv.size());
  size_t o_sg = 0, o_so = 0;
  if (use_split) {
    o_sg = put(sched.segs.data(), sizeof(HadiSegment) * sched.segs.size());
    o_so = put(sched.off.data(), sizeof(int) * sched.off.size());
  }
  cudaError_t e = cudaMemcpyAsync(d_stage, h_stage, off, cudaMemcpyHostToDevice, ctx->stream);
  ctx->h2d_bytes += (long long)off;
  if (e == cudaSuccess && use_dev_pool && !s_pool.empty()) {
    // grids seen for the first time: into the context's pool, ahead of the kernel on the same stream
    e = cudaMemcpyAsync(ctx->d_spool + ctx->spool_used, h_stage + o_s, sizeof(double) * s_pool.size(),
                        cudaMemcpyHostToDevice, ctx->stream);
    ctx->h2d_bytes += (long long)(sizeof(double) * s_pool.size());
  }
  if (e != cudaSuccess) {
    release_all();
    return cuda_fail(ctx, e, "H2D");
  }
  if (use_dev_pool) {
    ctx->spool_used += s_pool.size();
    ctx->s_dev.insert(pending.begin(), pending.end());
  }

  HadiLaunch& L = b->L;
  L.m1 = m1; L.m2 = m2; L.ld = g.ld; L.n1 = g.n1; L.n2 = g.n2; L.pj = g.pj;
  L.n_items = n_items;
  L.items = (const HadiItem*)(d_stage + o_items);
  L.s_pool = use_dev_pool ? ctx->d_spool : (const double*)(d_stage + o_s);
  L.v_pool = (const double*)(d_stage + o_v);
  L.e_pool = (const double*)(d_stage + o_e);
  L.nd = nd;
  L.div_dates = (const double*)(d_stage + o_d);
  L.div_amounts = L.div_dates + nd;
  L.div_pcts = L.div_dates + 2 * nd;
  L.scratch = d_scratch;
  L.scratch_stride = stride;
  L.counter = d_counter;
  L.out_values = d_values;
  L.out_stride = b->stride;
  L.reruns = reinterpret_cast<unsigned long long*>(d_values + (size_t)std::max(n_items, 1) * b->stride);
  L.out_U = nullptr;
  L.out_lam = nullptr;
  L.scheme = num->scheme;
  L.segs = nullptr; L.seg_off = nullptr; L.hand_state = nullptr; L.hand_data = nullptr;
  if (use_split) {
    int* hs = (int*)take(sizeof(int) * (size_t)sched.n_hand, false);
    double* hd = (double*)take(sizeof(double) * 2 * (size_t)P * (size_t)sched.n_hand, false);
    if (!hs || !hd) {
      cudaStreamSynchronize(ctx->stream);   // the staging copies above are in flight: do not hand their buffers back yet
      release_all();
      return HADI_ERR_NOMEM;
    }
    L.segs = (const HadiSegment*)(d_stage + o_sg);
    L.seg_off = (const int*)(d_stage + o_so);
    L.hand_state = hs;
    L.hand_data = hd;
    b->n_hand = sched.n_hand;
  }
  L.dbg_step = L.dbg_phase = 0;
  L.flags = HADI_FLAGS_DEFAULT;
  if (const char* fl = getenv("HADI_FLAGS")) L.flags = atoi(fl);
  if (const char* ds = getenv("HADI_DEBUG_STOP")) sscanf(ds, "%d:%d", &L.dbg_step, &L.dbg_phase);
  L.prof = (long long*)take(sizeof(long long) * (8 * (size_t)b->grid_ctas + 1), false);
  if (L.prof) cudaMemsetAsync(L.prof, 0, sizeof(long long) * (8 * (size_t)b->grid_ctas + 1), ctx->stream);
  if (cudaEventCreate(&b->ev0) != cudaSuccess || cudaEventCreate(&b->ev1) != cudaSuccess) {
    cudaStreamSynchronize(ctx->stream);
    if (b->ev0) cudaEventDestroy(b->ev0);
    b->ev0 = nullptr;
    release_all();
    return cuda_fail(ctx, cudaGetLastError(), "event");
  }
  *out = b.release();
  return HADI_OK;
} HADI_CATCH

int hadi_batch_num_items(const hadi_batch* b) { return b ? b->n_items : 0; }
long long hadi_batch_exact_reruns(const hadi_batch* b) { return b ? b->reruns : 0; }
int hadi_batch_values_per_item(const hadi_batch* b) { return b ? b->stride : 0; }
int hadi_batch_kernel_info(const hadi_batch* b, int* variant, int* grid_ctas, int* ctas_per_solve) try {
  if (!b) return HADI_ERR_ARG;
  if (variant) *variant = b->plan.variant;
  if (grid_ctas) *grid_ctas = b->grid_ctas;
  if (ctas_per_solve) *ctas_per_solve = std::max(1, b->plan.cluster);
  return HADI_OK;
} HADI_CATCH
double* hadi_batch_values_dev(hadi_batch* b) { return b ? b->L.out_values : nullptr; }

// Re-aim a prepared batch at new Heston parameters (kappa, eta, sigma, rho, V0) without rebuilding it: the LM loop
// solves the same options dozens of times, and strikes, step tables, schedule and buffers do not depend on the
// parameters.  Rewrites the parameter fields of the item descriptors and the (up to three) v-grids and uploads those
// two ranges; S0, r_d, r_f must be the ones the batch was created with.  The previous launch must have been fetched.
int hadi_batch_update_model(hadi_batch* b, const hadi_model* model) try {
  if (!b || !model) return HADI_ERR_ARG;
  hadi_ctx* ctx = b->ctx;
  if (bits(model->S0) != bits(b->model0.S0) || bits(model->r_d) != bits(b->model0.r_d) || bits(model->r_f) != bits(b->model0.r_f))
    return fail(ctx, HADI_ERR_ARG, "hadi_batch_update_model: S0, r_d and r_f are fixed at creation");
  cudaSetDevice(ctx->device);
  const int m2 = b->m2, nc = n_columns(b->mode);
  std::vector<double> v_pool((size_t)3 * (m2 + 1));
  v_grid(ctx, m2, model->V0, v_pool.data());
  const double V0p = model->V0 + b->eps5[4], V0m = model->V0 - b->eps5[4];
  v_grid(ctx, m2, V0p, v_pool.data() + (m2 + 1));
  v_grid(ctx, m2, b->mode == HADI_MODE_JACOBIAN_CENTRAL ? V0m : model->V0, v_pool.data() + 2 * (m2 + 1));
  int idx_v0 = find_node(v_pool.data(), m2 + 1, model->V0);
  if (idx_v0 < 0) idx_v0 = 0;
  int idx_v1 = find_node(v_pool.data() + (m2 + 1), m2 + 1, V0p);
  if (idx_v1 < 0) idx_v1 = 0;
  int idx_v2 = find_node(v_pool.data() + 2 * (m2 + 1), m2 + 1, V0m);
  if (idx_v2 < 0) idx_v2 = 0;
  int br_lo = 0, br_hi = 0;
  double br_w = 0.0;
  v0_bracket(v_pool.data(), m2, V0p, &br_lo, &br_hi, &br_w);
  for (HadiItem& it : b->h_items) {
    const int col = (b->item_begin + it.out) % nc;
    it.kappa = model->kappa;
    it.eta = model->eta;
    it.sigma = model->sigma;
    it.rho = model->rho;
    if (col == 1) it.kappa += b->eps5[0];
    if (col == 2) it.eta += b->eps5[1];
    if (col == 3) it.sigma += b->eps5[2];
    if (col == 4) it.rho += b->eps5[3];
    if (col == 6) it.kappa -= b->eps5[0];
    if (col == 7) it.eta -= b->eps5[1];
    if (col == 8) it.sigma -= b->eps5[2];
    if (col == 9) it.rho -= b->eps5[3];
    it.idx_v = (col == 5) ? idx_v1 : (col == 10) ? idx_v2 : idx_v0;
    it.aux = br_lo | (br_hi << 16);
  }
  const size_t bytes_items = sizeof(HadiItem) * b->h_items.size(), bytes_v = sizeof(double) * v_pool.size();
  if (bytes_items) std::memcpy(b->h_stage + b->o_items, b->h_items.data(), bytes_items);
  std::memcpy(b->h_stage + b->o_v, v_pool.data(), bytes_v);
  cudaError_t e = cudaSuccess;
  if (bytes_items) e = cudaMemcpyAsync(b->d_stage + b->o_items, b->h_stage + b->o_items, bytes_items, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(b->d_stage + b->o_v, b->h_stage + b->o_v, bytes_v, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "H2D (model update)");
  ctx->h2d_bytes += (long long)(bytes_items + bytes_v);
  return HADI_OK;
} HADI_CATCH

int hadi_batch_launch(hadi_batch* b) try {
  if (!b) return HADI_ERR_ARG;
  hadi_ctx* ctx = b->ctx;
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaMemsetAsync(b->L.counter, 0, sizeof(int) * HADI_COUNTER_INTS, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(b->L.reruns, 0, sizeof(unsigned long long), ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "memset");
  if (b->n_hand > 0) {
    e = cudaMemsetAsync(b->L.hand_state, 0, sizeof(int) * (size_t)b->n_hand, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "memset");
  }
  cudaEventRecord(b->ev0, ctx->stream);
  if (b->n_items > 0) {
    const int rc = b->plan.variant == HADI_WIDE_VARIANT ? hadi_launch_wide(b->L, b->plan, b->grid_ctas, ctx->stream)
                                                        : hadi_launch_douglas(b->L, b->plan, b->grid_ctas, ctx->stream);
    if (rc != 0) return cuda_fail(ctx, (cudaError_t)rc, "kernel launch");
    ctx->launches++;
  }
  cudaEventRecord(b->ev1, ctx->stream);
  b->launched = true;
  return HADI_OK;
} HADI_CATCH

int hadi_batch_fetch(hadi_batch* b, double* values) try {
  if (!b || (!values && b->n_items > 0)) return HADI_ERR_ARG;
  hadi_ctx* ctx = b->ctx;
  cudaSetDevice(ctx->device);
  const size_t nv = (size_t)b->n_items * (size_t)b->stride;
  const size_t slot = (size_t)std::max(b->n_items, 1) * (size_t)b->stride;   // where the re-solve counter sits
  cudaError_t e = cudaMemcpyAsync(b->h_values, b->L.out_values, sizeof(double) * (slot + 1),
                                  cudaMemcpyDeviceToHost, ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "D2H");
  ctx->d2h_bytes += (long long)(sizeof(double) * nv);
  e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "kernel execution");
  if (nv > 0) std::memcpy(values, b->h_values, sizeof(double) * nv);
  unsigned long long rr = 0;
  std::memcpy(&rr, b->h_values + slot, sizeof rr);
  b->reruns = (long long)rr;
  ctx->exact_reruns += (long long)rr;
  return HADI_OK;
} HADI_CATCH

int hadi_batch_elapsed_ms(hadi_batch* b, float* ms) try {
  if (!b || !ms || !b->launched) return HADI_ERR_ARG;
  cudaSetDevice(b->ctx->device);
  cudaError_t e = cudaEventSynchronize(b->ev1);
  if (e != cudaSuccess) return cuda_fail(b->ctx, e, "event sync");
  e = cudaEventElapsedTime(ms, b->ev0, b->ev1);
  if (e != cudaSuccess) return cuda_fail(b->ctx, e, "event elapsed");
  return HADI_OK;
} HADI_CATCH

// Phase cycle counters of the last launch summed over CTAs (all zero unless the library was built
// with -DHADI_PHASE_TIMING): [0] set-up, [1] dividend jump, [2] explicit stage, [3] A1 solves,
// [4] A2 solves, [5] projection, [6] output, [7] item fetch.  Development aid.
int hadi_batch_phase_cycles(hadi_batch* b, long long* out8) try {
  if (!b || !out8 || !b->L.prof) return HADI_ERR_ARG;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  std::vector<long long> h((size_t)8 * b->grid_ctas + 1);
  if (cudaMemcpy(h.data(), b->L.prof, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost) != cudaSuccess)
    return cuda_fail(b->ctx, cudaGetLastError(), "D2H prof");
  for (int k = 0; k < 8; ++k) out8[k] = 0;
  for (int c = 0; c < b->grid_ctas; ++c)
    for (int k = 0; k < 8; ++k) out8[k] += h[(size_t)c * 8 + k];
  return HADI_OK;
} HADI_CATCH

// development aid (not part of include/hadi.h): raw per-CTA debug records, 8 long long per CTA
int hadi_batch_prof_raw(hadi_batch* b, long long* out, int max_ctas) try {
  if (!b || !out || !b->L.prof) return HADI_ERR_ARG;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  const int n = std::min(max_ctas, b->grid_ctas);
  if (cudaMemcpy(out, b->L.prof, sizeof(long long) * 8 * (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess)
    return cuda_fail(b->ctx, cudaGetLastError(), "D2H prof");
  return n;
} HADI_CATCH

void hadi_batch_destroy(hadi_batch* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  for (int id : b->bufs) b->ctx->pool[id].in_use = false;
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  delete b;
}

// ------------------------------------------------------------------------------------------------
static const double kNoEps[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
static int values_per_item(int mode) { return mode == HADI_MODE_JACOBIAN_INTERP ? 3 : 1; }

// values[(end - begin) * values_per_item(mode)]
static int run_items(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                     const hadi_point* points, int mode, const double* eps5, int begin, int end, double* values,
                     double* U_out, double* lam_out, float* ms) {
  hadi_batch* b = nullptr;
  static const bool host_timing = getenv("HADI_HOST_TIMING") != nullptr;   // development aid: stage times on stderr
  const auto ht0 = std::chrono::steady_clock::now();
  int rc = hadi_batch_create_ex(ctx, model, num, n, points, mode, eps5, begin, end, &b);
  if (rc != HADI_OK) return rc;
  const auto ht1 = std::chrono::steady_clock::now();
  const size_t P = (size_t)(num->m1 + 1) * (size_t)(num->m2 + 1);
  int idU = -1, idL = -1;
  if (U_out) {
    idU = pool_get(ctx, sizeof(double) * P * (size_t)std::max(b->n_items, 1), false);
    if (idU < 0) { hadi_batch_destroy(b); return HADI_ERR_NOMEM; }
    b->bufs.push_back(idU);
    b->L.out_U = (double*)ctx->pool[idU].p;
  }
  if (lam_out) {
    idL = pool_get(ctx, sizeof(double) * P * (size_t)std::max(b->n_items, 1), false);
    if (idL < 0) { hadi_batch_destroy(b); return HADI_ERR_NOMEM; }
    b->bufs.push_back(idL);
    b->L.out_lam = (double*)ctx->pool[idL].p;
    cudaMemsetAsync(b->L.out_lam, 0, sizeof(double) * P * (size_t)std::max(b->n_items, 1), ctx->stream);
  }
  rc = hadi_batch_launch(b);
  if (rc == HADI_OK) rc = hadi_batch_fetch(b, values);
  if (rc == HADI_OK && U_out)
    if (cudaMemcpy(U_out, b->L.out_U, sizeof(double) * P * (size_t)b->n_items, cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = cuda_fail(ctx, cudaGetLastError(), "D2H U");
  if (rc == HADI_OK && lam_out)
    if (cudaMemcpy(lam_out, b->L.out_lam, sizeof(double) * P * (size_t)b->n_items, cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = cuda_fail(ctx, cudaGetLastError(), "D2H lambda");
  if (rc == HADI_OK && ms) hadi_batch_elapsed_ms(b, ms);
  if (host_timing && rc == HADI_OK) {
    const auto ht2 = std::chrono::steady_clock::now();
    float kms = 0.f;
    hadi_batch_elapsed_ms(b, &kms);
    fprintf(stderr, "[hadi] items %d: create %.1f us, launch+kernel+fetch %.1f us (kernel %.1f us)\n", b->n_items,
            std::chrono::duration<double, std::micro>(ht1 - ht0).count(),
            std::chrono::duration<double, std::micro>(ht2 - ht1).count(), kms * 1e3);
  }
  hadi_batch_destroy(b);
  return rc;
}

int hadi_price_batch(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                     const hadi_point* points, double* prices, double* U_out, double* lambda_out) try {
  if (!ctx || !prices) return HADI_ERR_ARG;
  if (n < 0 || (n > 0 && !points)) return fail(ctx, HADI_ERR_ARG, "bad argument");
  std::vector<double> vals((size_t)std::max(n, 1));
  const int rc = run_items(ctx, model, num, n, points, HADI_MODE_PRICE, kNoEps, 0, -1, vals.data(), U_out, lambda_out, nullptr);
  if (rc != HADI_OK) return rc;
  // results land at CalibrationPoint::global_index, as in the reference's multi-maturity drivers
  for (int k = 0; k < n; ++k) {
    const int gi = points[k].global_index;
    if (gi < 0 || gi >= n) return fail(ctx, HADI_ERR_ARG, "global_index out of range");
    prices[gi] = vals[k];
  }
  return HADI_OK;
} HADI_CATCH

int hadi_jacobian_batch(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                        const hadi_point* points, double eps, double* J, double* base_prices) try {
  hadi_jacobian_options jo{};
  jo.mode = HADI_MODE_JACOBIAN;
  for (int c = 0; c < 5; ++c) jo.eps[c] = eps;
  return hadi_jacobian_batch_ex(ctx, model, num, n, points, &jo, J, base_prices);
} HADI_CATCH

static bool valid_jopt(const hadi_jacobian_options* jo) {
  if (!jo) return false;
  if (jo->schedule != HADI_LM_SCHEDULE_REFERENCE && jo->schedule != HADI_LM_SCHEDULE_SPECULATIVE) return false;
  if (jo->mode != HADI_MODE_JACOBIAN && jo->mode != HADI_MODE_JACOBIAN_INTERP && jo->mode != HADI_MODE_JACOBIAN_CENTRAL)
    return false;
  for (int c = 0; c < 5; ++c)
    if (!(jo->eps[c] > 0.0)) return false;
  return true;
}

int hadi_jacobian_batch_ex(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                           const hadi_point* points, const hadi_jacobian_options* jo, double* J,
                           double* base_prices) try {
  if (!ctx || !J || !base_prices) return HADI_ERR_ARG;
  if (!model || n < 0 || (n > 0 && !points) || !valid_numerics(num) || !valid_jopt(jo)) return fail(ctx, HADI_ERR_ARG, "bad argument");
  const size_t per_option = (size_t)n_columns(jo->mode) * values_per_item(jo->mode);
  std::vector<double> vals(std::max<size_t>(per_option * n, 1)), Jt((size_t)std::max(5 * n, 1)), bt((size_t)std::max(n, 1));
  int rc = run_items(ctx, model, num, n, points, jo->mode, jo->eps, 0, -1, vals.data(), nullptr, nullptr, nullptr);
  if (rc != HADI_OK) return rc;
  int lo, hi;
  double w = 0.0;
  hadi_jacobian_v0_weight(num->m2, model->V0, jo->eps[4], &lo, &hi, &w);
  hadi_jacobian_assemble_ex(n, jo->mode, vals.data(), jo->eps, w, Jt.data(), bt.data());
  for (int k = 0; k < n; ++k) {
    const int gi = points[k].global_index;
    if (gi < 0 || gi >= n) return fail(ctx, HADI_ERR_ARG, "global_index out of range");
    base_prices[gi] = bt[k];
    for (int c = 0; c < 5; ++c) J[5 * gi + c] = Jt[5 * k + c];
  }
  return HADI_OK;
} HADI_CATCH

// One-call entry points over the attached communicator: every rank passes the same arguments, solves its
// cost-balanced slice of the work items and receives every result (SPMD).
static int solve_all(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                     const hadi_point* points, int mode, const double* eps5, const hadi_comm* comm, double* all,
                     float* ms);

int hadi_price_batch_sharded(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                             const hadi_point* points, double* prices) try {
  if (!ctx || !prices) return HADI_ERR_ARG;
  if (!model || n < 0 || (n > 0 && !points) || !valid_numerics(num)) return fail(ctx, HADI_ERR_ARG, "bad argument");
  std::vector<double> vals((size_t)std::max(n, 1));
  const int rc = solve_all(ctx, model, num, n, points, HADI_MODE_PRICE, kNoEps, nullptr, vals.data(), nullptr);
  if (rc != HADI_OK) return rc;
  for (int k = 0; k < n; ++k) {
    const int gi = points[k].global_index;
    if (gi < 0 || gi >= n) return fail(ctx, HADI_ERR_ARG, "global_index out of range");
    prices[gi] = vals[k];
  }
  return HADI_OK;
} HADI_CATCH

int hadi_jacobian_batch_sharded(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                                const hadi_point* points, const hadi_jacobian_options* jo, double* J,
                                double* base_prices) try {
  if (!ctx || !J || !base_prices) return HADI_ERR_ARG;
  if (!model || n < 0 || (n > 0 && !points) || !valid_numerics(num) || !valid_jopt(jo)) return fail(ctx, HADI_ERR_ARG, "bad argument");
  const size_t per_option = (size_t)n_columns(jo->mode) * values_per_item(jo->mode);
  std::vector<double> vals(std::max<size_t>(per_option * n, 1)), Jt((size_t)std::max(5 * n, 1)), bt((size_t)std::max(n, 1));
  int rc = solve_all(ctx, model, num, n, points, jo->mode, jo->eps, nullptr, vals.data(), nullptr);
  if (rc != HADI_OK) return rc;
  int lo, hi;
  double w = 0.0;
  hadi_jacobian_v0_weight(num->m2, model->V0, jo->eps[4], &lo, &hi, &w);
  hadi_jacobian_assemble_ex(n, jo->mode, vals.data(), jo->eps, w, Jt.data(), bt.data());
  for (int k = 0; k < n; ++k) {
    const int gi = points[k].global_index;
    if (gi < 0 || gi >= n) return fail(ctx, HADI_ERR_ARG, "global_index out of range");
    base_prices[gi] = bt[k];
    for (int c = 0; c < 5; ++c) J[5 * gi + c] = Jt[5 * k + c];
  }
  return HADI_OK;
} HADI_CATCH

// ---- Levenberg-Marquardt ----------------------------------------------------------------------
// src/jacobian_computation.cpp:20-104
int hadi_solve5(const double* Ain, const double* bin, double* x) try {
  if (!Ain || !bin || !x) return HADI_ERR_ARG;
  const int N = 5;
  double A[25], b[5];
  for (int i = 0; i < N; ++i) {
    b[i] = bin[i];
    for (int j = 0; j < N; ++j) A[i * N + j] = Ain[i * N + j];
  }
  for (int k = 0; k < N; ++k) {
    double maxA = std::fabs(A[k * N + k]);
    int piv = k;
    for (int p = k + 1; p < N; ++p) {
      const double val = std::fabs(A[p * N + k]);
      if (val > maxA) {
        maxA = val;
        piv = p;
      }
    }
    if (piv != k) {
      for (int c = 0; c < N; ++c) std::swap(A[k * N + c], A[piv * N + c]);
      std::swap(b[k], b[piv]);
    }
    const double pivot = A[k * N + k];
    for (int c = k + 1; c < N; ++c) A[k * N + c] /= pivot;
    b[k] /= pivot;
    A[k * N + k] = 1.0;
    for (int i = k + 1; i < N; ++i) {
      const double f = A[i * N + k];
      for (int c = k + 1; c < N; ++c) A[i * N + c] -= f * A[k * N + c];
      b[i] -= f * b[k];
      A[i * N + k] = 0.0;
    }
  }
  for (int k = N - 1; k >= 0; --k) {
    double val = b[k];
    for (int c = k + 1; c < N; ++c) val -= A[k * N + c] * b[c];
    b[k] = val;
  }
  for (int i = 0; i < N; ++i) x[i] = b[i];
  return HADI_OK;
} HADI_CATCH

// src/jacobian_computation.cpp:107-195.  The 5x5 normal equations are formed on the host in the
// oracle's (ascending-k) summation order: 30 dot products of length n are not GPU work.
int hadi_lm_update(int n, const double* J, const double* r, double lambda, double* delta) try {
  if (n < 0 || !J || !r || !delta) return HADI_ERR_ARG;
  double A[25], g[5];
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 5; ++j) {
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc += J[k * 5 + i] * J[k * 5 + j];
      A[i * 5 + j] = acc;
    }
  for (int i = 0; i < 5; ++i) A[i * 5 + i] *= (1.0 + lambda);
  for (int i = 0; i < 5; ++i) {
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc += J[k * 5 + i] * r[k];
    g[i] = acc;
  }
  return hadi_solve5(A, g, delta);
} HADI_CATCH

// A prepared batch plus what the exchange step needs: this rank's slice of the work items, the per-rank counts, and
// how the values come back — directly (one GPU), through the context's NCCL communicator (the kernel epilogue
// publishes into this rank's slot of the device gather buffer, one in-place ncclAllGather behind the kernel on the same
// stream, one device-to-host copy: no host hop in the exchange), or through the caller's hadi_comm hook.
struct ShardedBatch {
  hadi_batch* b = nullptr;
  int vpi = 1, world = 1, rank = 0;
  bool nccl = false;
  const hadi_comm* hook = nullptr;
  std::vector<int> counts, displs;
  size_t mx = 1;                 // longest slice in doubles (the all-gather's element count)
  std::vector<double> mine;      // hook path only
};

static void sharded_destroy(ShardedBatch* sb) {
  if (sb->b) hadi_batch_destroy(sb->b);
  sb->b = nullptr;
}

static int sharded_create(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                          const hadi_point* points, int mode, const double* eps5, const hadi_comm* comm,
                          ShardedBatch* sb) {
  const int nc = n_columns(mode);
  const int total = n * nc;
  sb->vpi = values_per_item(mode);
  sb->hook = (comm && comm->world > 1) ? comm : nullptr;
  sb->nccl = !sb->hook && ctx->nccl && ctx->nccl_world > 1;
  sb->world = sb->hook ? comm->world : sb->nccl ? ctx->nccl_world : 1;
  sb->rank = sb->hook ? comm->rank : sb->nccl ? ctx->nccl_rank : 0;
  int my_begin = 0, my_end = total;
  sb->counts.assign(sb->world, total * sb->vpi);
  sb->displs.assign(sb->world, 0);
  sb->mx = std::max<size_t>((size_t)total * sb->vpi, 1);
  if (sb->world > 1) {
    std::vector<int> costs((size_t)std::max(total, 1));
    hadi_item_costs(num, n, points, mode, costs.data());
    sb->mx = 1;
    for (int r = 0; r < sb->world; ++r) {
      int b, e;
      hadi_partition(total, costs.data(), sb->world, r, &b, &e);
      sb->displs[r] = b * sb->vpi;
      sb->counts[r] = (e - b) * sb->vpi;
      sb->mx = std::max(sb->mx, (size_t)sb->counts[r]);
      if (r == sb->rank) { my_begin = b; my_end = e; }
    }
  }
  if (sb->hook) sb->mine.resize((size_t)std::max(sb->counts[sb->rank], 1));
  if (sb->nccl) {
    cudaSetDevice(ctx->device);
    if (ctx->gather_cap < sb->mx) {
      cudaStreamSynchronize(ctx->stream);
      if (ctx->d_gather) cudaFree(ctx->d_gather);
      if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
      ctx->d_gather = nullptr; ctx->h_gather = nullptr; ctx->gather_cap = 0;
      const size_t cap = std::max<size_t>(sb->mx + sb->mx / 4, 4096);
      if (cudaMalloc(&ctx->d_gather, sizeof(double) * cap * sb->world) != cudaSuccess ||
          cudaMallocHost(&ctx->h_gather, sizeof(double) * cap * sb->world) != cudaSuccess)
        return cuda_fail(ctx, cudaGetLastError(), "gather buffers");
      ctx->gather_cap = cap;
    }
  }
  return hadi_batch_create_ex(ctx, model, num, n, points, mode, eps5, my_begin, my_end, &sb->b);
}

// One solver call on a prepared sharded batch: `model` != nullptr re-aims it first (hadi_batch_update_model).
// all[total items * values per item] on every rank.
static int sharded_solve(hadi_ctx* ctx, ShardedBatch* sb, const hadi_model* model, double* all, float* ms) {
  hadi_batch* b = sb->b;
  int rc = model ? hadi_batch_update_model(b, model) : HADI_OK;
  if (rc != HADI_OK) return rc;
  if (sb->world <= 1) {
    rc = hadi_batch_launch(b);
    if (rc == HADI_OK) rc = hadi_batch_fetch(b, all);
  } else if (sb->hook) {
    rc = hadi_batch_launch(b);
    if (rc == HADI_OK) rc = hadi_batch_fetch(b, sb->mine.data());
    if (rc == HADI_OK) {
      if (!sb->hook->allgather) return fail(ctx, HADI_ERR_COMM, "no allgather hook");
      if (sb->hook->allgather(sb->hook->user, sb->mine.data(), sb->counts[sb->rank], all, sb->counts.data(),
                              sb->displs.data(), sb->world) != 0)
        return fail(ctx, HADI_ERR_COMM, "allgather failed");
    }
  } else {
    const size_t mx = sb->mx;
    const int W = sb->world, R = sb->rank;
    if (ctx->gather_cap < mx) return fail(ctx, HADI_ERR_COMM, "gather buffer smaller than the batch (context shared between calibrations?)");
    b->L.out_values = ctx->d_gather + (size_t)R * mx;   // the kernel epilogue publishes into the gather buffer
    rc = hadi_batch_launch(b);
    if (rc == HADI_OK) {
      const ncclResult_t nr = nccl_api().AllGather(ctx->d_gather + (size_t)R * mx, ctx->d_gather, mx, ncclDouble,
                                                   ctx->nccl, ctx->stream);
      if (nr != ncclSuccess) rc = nccl_fail(ctx, nr, "ncclAllGather");
    }
    if (rc == HADI_OK) {
      cudaError_t e = cudaMemcpyAsync(ctx->h_gather, ctx->d_gather, sizeof(double) * mx * W, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(b->h_values, b->L.reruns, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) rc = cuda_fail(ctx, e, "gather D2H");
      ctx->d2h_bytes += (long long)(sizeof(double) * mx * W);
    }
    if (rc == HADI_OK) {
      for (int r = 0; r < W; ++r)
        if (sb->counts[r] > 0)
          std::memcpy(all + sb->displs[r], ctx->h_gather + (size_t)r * mx, sizeof(double) * (size_t)sb->counts[r]);
      unsigned long long rr = 0;
      std::memcpy(&rr, b->h_values, sizeof rr);
      b->reruns = (long long)rr;
      ctx->exact_reruns += (long long)rr;
    }
  }
  if (rc == HADI_OK && ms) hadi_batch_elapsed_ms(b, ms);
  return rc;
}

// Solve the items of [0, n*nc) across the ranks (the context's communicator, or `comm`) and return every item value
// on every rank (all[n * nc * values_per_item(mode)]).
static int solve_all(hadi_ctx* ctx, const hadi_model* model, const hadi_numerics* num, int n,
                     const hadi_point* points, int mode, const double* eps5, const hadi_comm* comm, double* all,
                     float* ms) {
  ShardedBatch sb;
  int rc = sharded_create(ctx, model, num, n, points, mode, eps5, comm, &sb);
  if (rc == HADI_OK) rc = sharded_solve(ctx, &sb, nullptr, all, ms);
  sharded_destroy(&sb);
  return rc;
}

// src/heston_calibration.cpp:2692-2831 (multi-maturity; the single-maturity twin at :204-417 is the
// same loop with one (N, dt)).
int hadi_calibrate(hadi_ctx* ctx, const hadi_model* initial, const hadi_numerics* num, int n,
                   const hadi_point* points, const double* market, const hadi_lm_options* opt,
                   const hadi_comm* comm, hadi_lm_result* res) try {
  if (!opt) return HADI_ERR_ARG;
  hadi_jacobian_options jo{};
  jo.mode = HADI_MODE_JACOBIAN;
  for (int c = 0; c < 5; ++c) jo.eps[c] = opt->eps;
  return hadi_calibrate_ex(ctx, initial, num, n, points, market, opt, &jo, comm, res);
} HADI_CATCH

// The same loop with the Jacobian taken as `jo` says (opt->eps is ignored).
// HADI_MODE_JACOBIAN reproduces the reference's trajectory; the other modes change the numbers.
int hadi_calibrate_ex(hadi_ctx* ctx, const hadi_model* initial, const hadi_numerics* num, int n,
                      const hadi_point* points, const double* market, const hadi_lm_options* opt,
                      const hadi_jacobian_options* jo, const hadi_comm* comm, hadi_lm_result* res) try {
  if (!ctx || !initial || !points || !market || !opt || !res || n <= 0) return HADI_ERR_ARG;
  if (!valid_numerics(num) || !valid_jopt(jo)) return fail(ctx, HADI_ERR_ARG, "bad argument");
  const int jcols = n_columns(jo->mode);
  for (int k = 0; k < n; ++k)
    if (points[k].global_index < 0 || points[k].global_index >= n) return fail(ctx, HADI_ERR_ARG, "global_index out of range");
  hadi_model cur = *initial;
  const long long reruns0 = ctx->exact_reruns;
  double lambda = opt->lambda0;
  std::vector<double> vals((size_t)jcols * values_per_item(jo->mode) * n), J((size_t)5 * n), Jt((size_t)5 * n), base(n), bt(n), r(n), newp(n), nv(n);
  bool converged = false;
  int iters = 0, solves = 0;
  double final_error = 100.0, delta_norm = 0.0, gpu_ms = 0.0;
  // the two batches of the loop (Jacobian, candidate prices) are built once; every iteration only re-aims them at
  // the current parameters (hadi_batch_update_model)
  ShardedBatch sb_jac, sb_price;
  struct Guard {
    ShardedBatch *a, *b;
    ~Guard() { sharded_destroy(a); sharded_destroy(b); }
  } guard{&sb_jac, &sb_price};
  int rc = sharded_create(ctx, &cur, num, n, points, jo->mode, jo->eps, comm, &sb_jac);
  if (rc == HADI_OK) rc = sharded_create(ctx, &cur, num, n, points, HADI_MODE_PRICE, kNoEps, comm, &sb_price);
  if (rc != HADI_OK) return rc;
  const bool speculative = jo->schedule == HADI_LM_SCHEDULE_SPECULATIVE;
  std::vector<double> cand_vals(speculative ? vals.size() : 0);
  bool have_jacobian = false;   // speculative schedule: vals already holds the item values at `cur`
  for (int iter = 0; iter < opt->max_iter && !converged; ++iter) {
    float ms = 0.f;
    if (!have_jacobian) {
      rc = sharded_solve(ctx, &sb_jac, iter == 0 ? nullptr : &cur, vals.data(), &ms);
      if (rc != HADI_OK) return rc;
      gpu_ms += ms;
      solves += jcols * n;
    }
    int blo, bhi;
    double bw = 0.0;
    hadi_jacobian_v0_weight(num->m2, cur.V0, jo->eps[4], &blo, &bhi, &bw);
    hadi_jacobian_assemble_ex(n, jo->mode, vals.data(), jo->eps, bw, Jt.data(), bt.data());
    for (int k = 0; k < n; ++k) {
      const int gi = points[k].global_index;
      base[gi] = bt[k];
      for (int c = 0; c < 5; ++c) J[5 * gi + c] = Jt[5 * k + c];
    }
    for (int i = 0; i < n; ++i) r[i] = market[i] - base[i];
    double delta[5];
    hadi_lm_update(n, J.data(), r.data(), lambda, delta);
    hadi_model nw = cur;
    nw.kappa = std::max(1e-3, cur.kappa + delta[0]);
    nw.eta = std::max(1e-2, cur.eta + delta[1]);
    nw.sigma = std::max(1e-2, cur.sigma + delta[2]);
    nw.rho = std::min(1.0, std::max(-1.0, cur.rho + delta[3]));
    nw.V0 = std::max(1e-2, cur.V0 + delta[4]);
    delta_norm = 0.0;
    for (int i = 0; i < 5; ++i) delta_norm += delta[i] * delta[i];
    delta_norm = std::sqrt(delta_norm);
    double cur_err = 0;
    for (int i = 0; i < n; ++i) cur_err += r[i] * r[i];
    if (delta_norm < opt->delta_tol || cur_err < opt->tol) {
      converged = true;
      cur = nw;  // the candidate is accepted without being evaluated (:2763-2780)
      final_error = cur_err;
      iters = iter + 1;
      break;
    }
    if (speculative) {
      // the candidate's prices are the base column of ITS Jacobian batch: solve that, and keep it if the step is accepted
      rc = sharded_solve(ctx, &sb_jac, &nw, cand_vals.data(), &ms);
      if (rc != HADI_OK) return rc;
      gpu_ms += ms;
      solves += jcols * n;
      const size_t per_option = (size_t)jcols * values_per_item(jo->mode);
      for (int k = 0; k < n; ++k) newp[points[k].global_index] = cand_vals[per_option * k];
    } else {
      rc = sharded_solve(ctx, &sb_price, &nw, nv.data(), &ms);
      if (rc != HADI_OK) return rc;
      gpu_ms += ms;
      solves += n;
      for (int k = 0; k < n; ++k) newp[points[k].global_index] = nv[k];
    }
    double new_err = 0.0;
    for (int i = 0; i < n; ++i) {
      const double rr = market[i] - newp[i];
      new_err += rr * rr;
    }
    if (new_err < cur_err) {
      cur = nw;
      lambda = std::max(lambda / 10.0, 1e-7);
      if (speculative) vals.swap(cand_vals);   // the Jacobian of the next iteration is already here
    } else {
      lambda = std::min(lambda * 10.0, 1e7);
    }
    have_jacobian = speculative;   // accepted: the candidate's batch; rejected: the Jacobian at `cur` is still valid
    final_error = std::min(new_err, cur_err);
    iters = iter + 1;
  }
  res->params[0] = cur.kappa;
  res->params[1] = cur.eta;
  res->params[2] = cur.sigma;
  res->params[3] = cur.rho;
  res->params[4] = cur.V0;
  res->final_error = final_error;
  res->lambda = lambda;
  res->delta_norm = delta_norm;
  res->iterations = iters;
  res->converged = converged ? 1 : 0;
  res->pde_solves = solves;
  res->gpu_ms = gpu_ms;
  res->exact_reruns = ctx->exact_reruns - reruns0;
  return HADI_OK;
} HADI_CATCH

}  // extern "C"
