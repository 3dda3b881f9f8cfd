// capi.cu -- the extern "C" surface declared in include/mrs_b200.h (engine, model queries, fit / MAE /
// predict wrappers).  Kernels live in loader.cu, baseline.cu and knn.cu.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace mrs {

static thread_local char g_err[1024] = "";
thread_local mrs_engine* tls_engine = nullptr;

constexpr size_t kCacheCap = (size_t)64 << 30;  // keep at most this many bytes of released buffers per engine

void* cache_alloc(size_t bytes) {
  bytes = (bytes + 255) & ~(size_t)255;
  mrs_engine* e = tls_engine;
  if (e) {
    auto it = e->free_blocks.find(bytes);
    if (it != e->free_blocks.end()) {
      void* p = it->second;
      e->free_blocks.erase(it);
      e->cached_bytes -= bytes;
      e->live_blocks[p] = bytes;
      return p;
    }
  }
  void* p = nullptr;
  cudaError_t err = cudaMalloc(&p, bytes);
  if (err != cudaSuccess && e && !e->free_blocks.empty()) {  // give the cache back to the driver and retry
    cudaStreamSynchronize(e->stream);
    for (auto& kv : e->free_blocks) cudaFree(kv.second);
    e->free_blocks.clear();
    e->cached_bytes = 0;
    err = cudaMalloc(&p, bytes);
  }
  if (err != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(err));
    return nullptr;
  }
  if (e) e->live_blocks[p] = bytes;
  return p;
}

void cache_free(void* p) {
  mrs_engine* e = tls_engine;
  if (e) {
    auto it = e->live_blocks.find(p);
    if (it != e->live_blocks.end()) {
      const size_t bytes = it->second;
      e->live_blocks.erase(it);
      if (e->cached_bytes + bytes <= kCacheCap) {
        e->free_blocks.emplace(bytes, p);
        e->cached_bytes += bytes;
        return;
      }
    }
  }
  cudaFree(p);
}
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int32_t ensure_scratch(mrs_engine* e, size_t bytes) {
  if (bytes <= e->scratch_bytes) return MRS_OK;
  if (e->scratch) {
    cudaStreamSynchronize(e->stream);
    cudaFree(e->scratch);
    e->scratch = nullptr;
    e->scratch_bytes = 0;
  }
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t err = cudaMalloc(&e->scratch, want);
  if (err != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes of scratch) failed: %s", want, cudaGetErrorString(err));
    return MRS_ERR_NOMEM;
  }
  e->scratch_bytes = want;
  return MRS_OK;
}

std::vector<int3> deal_ctas(const std::vector<int64_t>& tile_cost, int32_t n_ctas) {
  const int32_t nt = (int32_t)tile_cost.size();
  int64_t total = 0;
  int32_t busy = 0;
  for (int64_t c : tile_cost) { total += c; busy += c > 0; }
  std::vector<int3> desc;
  if (total == 0 || n_ctas <= 0) return desc;
  std::vector<int32_t> share((size_t)nt, 0);
  if (busy >= n_ctas) {  // more busy tiles than CTAs: one CTA per busy tile (the grid is then larger than one wave)
    for (int32_t t = 0; t < nt; ++t) share[(size_t)t] = tile_cost[(size_t)t] > 0 ? 1 : 0;
  } else {
    // one CTA per busy tile, the rest by largest remainder of the exact proportional share
    std::vector<double> want((size_t)nt, 0.0);
    int32_t given = 0;
    for (int32_t t = 0; t < nt; ++t)
      if (tile_cost[(size_t)t] > 0) {
        want[(size_t)t] = (double)tile_cost[(size_t)t] * (double)n_ctas / (double)total;
        share[(size_t)t] = std::max<int32_t>(1, (int32_t)want[(size_t)t]);
        given += share[(size_t)t];
      }
    while (given < n_ctas) {  // hand the remaining CTAs to the tiles with the most cost per CTA
      int32_t best = -1;
      double worst = -1.0;
      for (int32_t t = 0; t < nt; ++t)
        if (share[(size_t)t] > 0 && (double)tile_cost[(size_t)t] / share[(size_t)t] > worst) { worst = (double)tile_cost[(size_t)t] / share[(size_t)t]; best = t; }
      ++share[(size_t)best];
      ++given;
    }
    while (given > n_ctas) {  // (the floor of one CTA per busy tile can overshoot): take from the tiles with the least cost per CTA
      int32_t best = -1;
      double least = 1e300;
      for (int32_t t = 0; t < nt; ++t)
        if (share[(size_t)t] > 1 && (double)tile_cost[(size_t)t] / (share[(size_t)t] - 1) < least) { least = (double)tile_cost[(size_t)t] / (share[(size_t)t] - 1); best = t; }
      if (best < 0) break;
      --share[(size_t)best];
      --given;
    }
  }
  for (int32_t t = 0; t < nt; ++t)
    for (int32_t k = 0; k < share[(size_t)t]; ++k) desc.push_back(make_int3(t, k, share[(size_t)t]));
  return desc;
}

}  // namespace mrs

using namespace mrs;

extern "C" const char* mrs_last_error(void) { return g_err; }
extern "C" const char* mrs_version(void) { return "mrs_b200 0.1.0 (sm_100a)"; }
extern "C" int64_t mrs_launch_count(void) { return (int64_t)g_launches.load(); }

// ------------------------------------------------------------------ engine
extern "C" int32_t mrs_engine_create(int32_t device, void* cuda_stream, mrs_engine** out) {
  MRS_REQUIRE(out, MRS_ERR_INVALID, "mrs_engine_create: NULL output");
  int n_dev = 0;
  cudaError_t ce = cudaGetDeviceCount(&n_dev);
  if (ce != cudaSuccess || n_dev == 0) {
    // no CPU fallback by design: the engine only exists on a CUDA device
    set_error("mrs_engine_create: no CUDA device available (%s)", ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0");
    return MRS_ERR_CUDA;
  }
  MRS_REQUIRE(device >= 0 && device < n_dev, MRS_ERR_INVALID, "mrs_engine_create: device %d out of range [0,%d)", device, n_dev);
  MRS_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MRS_CUDA(cudaGetDeviceProperties(&prop, device));
  MRS_REQUIRE(prop.major >= 10, MRS_ERR_UNSUPPORTED, "mrs_engine_create: device %d is sm_%d%d; this library is built for sm_100a only",
              device, prop.major, prop.minor);
  mrs_engine* e = new mrs_engine();
  e->device = device;
  e->sm_count = prop.multiProcessorCount;
  if (cuda_stream) {
    e->stream = (cudaStream_t)cuda_stream;
    e->own_stream = false;
  } else {
    cudaError_t se = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (se != cudaSuccess) { delete e; set_error("cudaStreamCreate failed: %s", cudaGetErrorString(se)); return MRS_ERR_CUDA; }
    e->own_stream = true;
  }
  use_engine(e);
  if (cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_order, cudaEventDisableTiming) != cudaSuccess) {
    set_error("mrs_engine_create: copy stream / event creation failed");
    mrs_engine_destroy(e);
    return MRS_ERR_CUDA;
  }
  if (getenv("MRS_TIMELINE") && atoi(getenv("MRS_TIMELINE"))) {  // diagnostics: per-kernel start/end stamps of a pass
    if (cudaMalloc((void**)&e->d_timeline, (32 + 1024 + 16384) * sizeof(unsigned long long)) != cudaSuccess) e->d_timeline = nullptr;
    if (e->d_timeline) cudaMemset(e->d_timeline, 0, (32 + 1024 + 16384) * sizeof(unsigned long long));
  }
  cudaError_t he = cudaMallocHost((void**)&e->h_pinned, 64 * sizeof(double));
  if (he != cudaSuccess) { set_error("cudaMallocHost failed: %s", cudaGetErrorString(he)); mrs_engine_destroy(e); return MRS_ERR_NOMEM; }
  *out = e;
  return MRS_OK;
}

extern "C" void mrs_engine_destroy(mrs_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->copy_stream) { cudaStreamSynchronize(e->copy_stream); cudaStreamDestroy(e->copy_stream); }
  if (e->ev_order) cudaEventDestroy(e->ev_order);
  if (e->scratch) cudaFree(e->scratch);
  if (e->d_timeline) cudaFree(e->d_timeline);
  for (auto& kv : e->free_blocks) cudaFree(kv.second);
  e->free_blocks.clear();
  if (tls_engine == e) tls_engine = nullptr;
  if (e->h_pinned) cudaFreeHost(e->h_pinned);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

extern "C" int32_t mrs_debug_timeline(mrs_engine* e, uint64_t* out32) {
  MRS_REQUIRE(e && out32, MRS_ERR_INVALID, "mrs_debug_timeline: NULL argument");
  MRS_REQUIRE(e->d_timeline, MRS_ERR_INVALID, "mrs_debug_timeline: the engine was created without MRS_TIMELINE=1");
  MRS_CUDA(cudaStreamSynchronize(e->stream));
  MRS_CUDA(cudaMemcpy(out32, e->d_timeline, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  uint64_t init[32];
  for (int k = 0; k < 32; ++k) init[k] = (k & 1) ? 0ull : ~0ull;
  MRS_CUDA(cudaMemcpy(e->d_timeline, init, sizeof(init), cudaMemcpyHostToDevice));
  return MRS_OK;
}

// per-CTA stamps of the last pass (MRS_TIMELINE=1): [0,256) item pass: table built, streaming starts; [256,512) item pass: CTA done;
// [512,768) test pass: table built; [768,1024) test pass: CTA done (ns of %globaltimer; 0 = CTA index not used)
extern "C" int32_t mrs_debug_cta_stamps(mrs_engine* e, uint64_t* out1024) {
  MRS_REQUIRE(e && out1024, MRS_ERR_INVALID, "mrs_debug_cta_stamps: NULL argument");
  MRS_REQUIRE(e->d_timeline, MRS_ERR_INVALID, "mrs_debug_cta_stamps: the engine was created without MRS_TIMELINE=1");
  MRS_CUDA(cudaStreamSynchronize(e->stream));
  MRS_CUDA(cudaMemcpy(out1024, e->d_timeline + 32, 1024 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  MRS_CUDA(cudaMemset(e->d_timeline + 32, 0, 1024 * sizeof(uint64_t)));
  return MRS_OK;
}
// per-warp stamps of the item pass (library built with -DMRS_WARP_STAMPS): [0,8192) end time of warp (cta*32 + w), [8192,16384) rows << 32 | slices
extern "C" int32_t mrs_debug_warp_stamps(mrs_engine* e, uint64_t* out16384) {
  MRS_REQUIRE(e && out16384, MRS_ERR_INVALID, "mrs_debug_warp_stamps: NULL argument");
  MRS_REQUIRE(e->d_timeline, MRS_ERR_INVALID, "mrs_debug_warp_stamps: the engine was created without MRS_TIMELINE=1");
  MRS_CUDA(cudaStreamSynchronize(e->stream));
  MRS_CUDA(cudaMemcpy(out16384, e->d_timeline + 32 + 1024, 16384 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return MRS_OK;
}

extern "C" int32_t mrs_engine_sync(mrs_engine* e) {
  MRS_REQUIRE(e, MRS_ERR_INVALID, "mrs_engine_sync: NULL engine");
  use_engine(e);
  MRS_CUDA(cudaStreamSynchronize(e->stream));
  return MRS_OK;
}

struct mrs_graph {
  mrs_engine* eng = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
};

extern "C" int32_t mrs_graph_begin(mrs_engine* e) {
  MRS_REQUIRE(e, MRS_ERR_INVALID, "mrs_graph_begin: NULL engine");
  MRS_REQUIRE(!e->profiling, MRS_ERR_INVALID, "mrs_graph_begin: per-kernel profiling is active");
  MRS_CUDA(cudaSetDevice(e->device));
  MRS_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
  return MRS_OK;
}

extern "C" int32_t mrs_graph_end(mrs_engine* e, mrs_graph** out) {
  MRS_REQUIRE(e && out, MRS_ERR_INVALID, "mrs_graph_end: NULL argument");
  *out = nullptr;
  cudaGraph_t g = nullptr;
  MRS_CUDA(cudaStreamEndCapture(e->stream, &g));
  cudaGraphExec_t ex = nullptr;
  cudaError_t ce = cudaGraphInstantiate(&ex, g, 0);
  if (ce != cudaSuccess) {
    cudaGraphDestroy(g);
    set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
    return MRS_ERR_CUDA;
  }
  mrs_graph* h = new mrs_graph();
  h->eng = e; h->graph = g; h->exec = ex;
  *out = h;
  return MRS_OK;
}

extern "C" int32_t mrs_graph_launch(mrs_graph* g) {
  MRS_REQUIRE(g, MRS_ERR_INVALID, "mrs_graph_launch: NULL graph");
  use_engine(g->eng);
  MRS_CUDA(cudaGraphLaunch(g->exec, g->eng->stream));
  return MRS_OK;
}

extern "C" void mrs_graph_destroy(mrs_graph* g) {
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
}

extern "C" int32_t mrs_profile_begin(mrs_engine* e) {
  MRS_REQUIRE(e, MRS_ERR_INVALID, "mrs_profile_begin: NULL engine");
  for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
  e->prof_events.clear();
  e->prof_names.clear();
  e->profiling = true;
  cudaEvent_t ev;
  MRS_CUDA(cudaEventCreate(&ev));
  MRS_CUDA(cudaEventRecord(ev, e->stream));
  e->prof_events.push_back(ev);
  e->prof_names.push_back("begin");
  return MRS_OK;
}

extern "C" int32_t mrs_profile_end(mrs_engine* e, char* names_out, int64_t names_cap, float* ms_out, int32_t cap, int32_t* n_out) {
  MRS_REQUIRE(e && n_out, MRS_ERR_INVALID, "mrs_profile_end: NULL argument");
  e->profiling = false;
  MRS_CUDA(cudaStreamSynchronize(e->stream));
  const int32_t n = (int32_t)e->prof_events.size() - 1;
  *n_out = n < 0 ? 0 : n;
  std::string names;
  for (int32_t j = 0; j < n; ++j) {
    float ms = 0.f;
    MRS_CUDA(cudaEventElapsedTime(&ms, e->prof_events[j], e->prof_events[j + 1]));
    if (ms_out && j < cap) ms_out[j] = ms;
    names += e->prof_names[j + 1];
    names += '\n';
  }
  if (names_out && names_cap > 0) {
    snprintf(names_out, (size_t)names_cap, "%s", names.c_str());
  }
  for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
  e->prof_events.clear();
  e->prof_names.clear();
  return MRS_OK;
}

// ------------------------------------------------------------------ fit
extern "C" int32_t mrs_fit_local(mrs_engine* e, const mrs_ratings* train, mrs_model** inout) { return fit_local(e, train, inout, false); }
extern "C" int32_t mrs_fit_async(mrs_engine* e, const mrs_ratings* train, mrs_model** inout) { return fit_local(e, train, inout, true); }
extern "C" int32_t mrs_fit_users_async(mrs_engine* e, const mrs_ratings* train, mrs_model** inout) { return fit_users(e, train, inout); }
extern "C" int32_t mrs_fit_finish(mrs_model* m) { return fit_finish(m); }

extern "C" int32_t mrs_fit(mrs_engine* e, const mrs_ratings* train, mrs_model** out) {
  MRS_REQUIRE(out, MRS_ERR_INVALID, "mrs_fit: NULL output");
  *out = nullptr;
  int32_t s = fit_local(e, train, out, true);
  if (s != MRS_OK) { mrs_model_destroy(*out); *out = nullptr; return s; }
  if (s == MRS_OK) {
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) { set_error("mrs_fit: %s", cudaGetErrorString(ce)); s = MRS_ERR_CUDA; }
  }
  if (s != MRS_OK) { mrs_model_destroy(*out); *out = nullptr; }
  return s;
}

extern "C" int32_t mrs_model_set_item_averages(mrs_model* m, int32_t enabled) {
  MRS_REQUIRE(m, MRS_ERR_INVALID, "mrs_model_set_item_averages: NULL model");
  m->want_item_avg = enabled != 0;
  m->finished = false;  // the next query needs a fit made with the new setting
  m->host_valid = false;
  return MRS_OK;
}

extern "C" int32_t mrs_model_exchange_buffer(mrs_model* m, void** device_ptr, int64_t* n_doubles) {
  MRS_REQUIRE(m && device_ptr && n_doubles, MRS_ERR_INVALID, "mrs_model_exchange_buffer: NULL argument");
  *device_ptr = m->xbuf;
  *n_doubles = 3 * (int64_t)m->n_items + 2;
  return MRS_OK;
}

extern "C" void mrs_model_destroy(mrs_model* m) {
  if (!m) return;
  if (m->eng) use_engine(m->eng);
  dev_free(m->uinv_hi); dev_free(m->uinv_lo);
  dev_free(m->usum); dev_free(m->k1_part); dev_free(m->xdev_fix); dev_free(m->xcode_sum);
  dev_free(m->upart); dev_free(m->uavg); dev_free(m->ipart); dev_free(m->xbuf); dev_free(m->idevavg); dev_free(m->iavg);
  dev_free(m->gavg); dev_free(m->mae_part); dev_free(m->counters); dev_free(m->slot_of_item); dev_free(m->item_slot); dev_free(m->tie_rank); dev_free(m->tie_inv);
  delete m;
}

// ------------------------------------------------------------------ model queries
static int32_t model_gavg(const mrs_model* m, double* out) {
  use_engine(m->eng);
  MRS_REQUIRE(m->finished, MRS_ERR_INVALID, "model not finished (call mrs_fit_finish)");
  if (!m->host_valid) {
    MRS_CUDA(cudaMemcpyAsync(&m->h_gavg, m->gavg, sizeof(double), cudaMemcpyDeviceToHost, m->eng->stream));
    MRS_CUDA(cudaStreamSynchronize(m->eng->stream));
    m->host_valid = true;
  }
  *out = m->h_gavg;
  return MRS_OK;
}

extern "C" int32_t mrs_model_scalar(const mrs_model* m, int32_t kind, double* out) {
  MRS_REQUIRE(m && out, MRS_ERR_INVALID, "mrs_model_scalar: NULL argument");
  MRS_REQUIRE(kind == MRS_GLOBAL_AVG, MRS_ERR_INVALID, "mrs_model_scalar: kind %d is not a scalar", kind);
  return model_gavg(m, out);
}

// copies `count` elements starting at `first` of a per-id table, plus the matching counts
static int32_t fetch_table(const mrs_model* m, int32_t kind, int64_t first, int64_t count, double* vals, int32_t* counts) {
  use_engine(m->eng);
  const mrs_ratings* R = m->train;
  cudaStream_t st = m->eng->stream;
  const double* src = nullptr;
  const int32_t* ptr = nullptr;
  switch (kind) {
    case MRS_USER_AVG: src = m->uavg; ptr = R->urow; break;
    case MRS_ITEM_AVG:
      MRS_REQUIRE(m->want_item_avg, MRS_ERR_INVALID, "item averages were switched off for this model (mrs_model_set_item_averages)");
      src = m->iavg; ptr = R->icolp; break;
    case MRS_ITEM_AVG_DEV: src = m->idevavg; ptr = R->icolp; break;
    default: set_error("unknown vector kind %d", kind); return MRS_ERR_INVALID;
  }
  std::vector<int32_t> p((size_t)count + 1);
  std::vector<double> xcnt;
  MRS_CUDA(cudaMemcpyAsync(vals, src + first, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
  const bool item_kind = (kind != MRS_USER_AVG);
  if (item_kind) {
    // counts of items come from the (possibly all-reduced) exchange buffer, not the local column pointer
    xcnt.resize((size_t)count);
    MRS_CUDA(cudaMemcpyAsync(xcnt.data(), m->xbuf + (size_t)m->n_items + first, sizeof(double) * count, cudaMemcpyDeviceToHost, st));
  } else {
    MRS_CUDA(cudaMemcpyAsync(p.data(), ptr + first, sizeof(int32_t) * (count + 1), cudaMemcpyDeviceToHost, st));
  }
  MRS_CUDA(cudaStreamSynchronize(st));
  double g = 0.0;
  MRS_TRY(model_gavg(m, &g));
  for (int64_t j = 0; j < count; ++j) {
    int32_t c = item_kind ? (int32_t)xcnt[j] : p[j + 1] - p[j];
    if (counts) counts[j] = c;
    if (c == 0) vals[j] = (kind == MRS_ITEM_AVG_DEV) ? 0.0 : g;  // SURVEY A.3 fallbacks
  }
  return MRS_OK;
}

extern "C" int32_t mrs_model_lookup(const mrs_model* m, int32_t kind, int32_t id, double* out, int32_t* known_out) {
  MRS_REQUIRE(m && out, MRS_ERR_INVALID, "mrs_model_lookup: NULL argument");
  MRS_REQUIRE(m->finished, MRS_ERR_INVALID, "mrs_model_lookup: model not finished");
  if (kind == MRS_GLOBAL_AVG) { if (known_out) *known_out = 1; return model_gavg(m, out); }
  const int32_t dim = (kind == MRS_USER_AVG) ? m->n_users : m->n_items;
  MRS_REQUIRE(kind == MRS_USER_AVG || kind == MRS_ITEM_AVG || kind == MRS_ITEM_AVG_DEV, MRS_ERR_INVALID, "mrs_model_lookup: unknown kind %d", kind);
  if (id < 0 || id >= dim) {
    if (known_out) *known_out = 0;
    if (kind == MRS_ITEM_AVG_DEV) { *out = 0.0; return MRS_OK; }
    return model_gavg(m, out);
  }
  int32_t c = 0;
  MRS_TRY(fetch_table(m, kind, id, 1, out, &c));
  if (known_out) *known_out = c > 0;
  return MRS_OK;
}

extern "C" int32_t mrs_model_vector(const mrs_model* m, int32_t kind, double* vals_out, int32_t* counts_out, int64_t cap, int64_t* n_out) {
  MRS_REQUIRE(m && n_out, MRS_ERR_INVALID, "mrs_model_vector: NULL argument");
  MRS_REQUIRE(m->finished, MRS_ERR_INVALID, "mrs_model_vector: model not finished");
  MRS_REQUIRE(kind == MRS_USER_AVG || kind == MRS_ITEM_AVG || kind == MRS_ITEM_AVG_DEV, MRS_ERR_INVALID, "mrs_model_vector: unknown kind %d", kind);
  const int64_t dim = (kind == MRS_USER_AVG) ? m->n_users : m->n_items;
  *n_out = dim;
  if (!vals_out) return MRS_OK;  // size query
  MRS_REQUIRE(cap >= dim, MRS_ERR_INVALID, "mrs_model_vector: capacity %lld < %lld", (long long)cap, (long long)dim);
  return fetch_table(m, kind, 0, dim, vals_out, counts_out);
}
