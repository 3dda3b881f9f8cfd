// query.cu -- batched prediction, fused MAE and top-n recommendation entry points of the C ABI.
#include <climits>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace mrs {
namespace {

__device__ __forceinline__ bool before(double ka, int32_t ia, double kb, int32_t ib) {
  return (ka > kb) || (ka == kb && ia < ib);  // P:654-660: score desc, then item id asc
}

// candidates of recommendations (P:667): items that occur in the train set and were not rated by `user`
__global__ void reco_pairs_kernel(int32_t user, int32_t n_items, int32_t* __restrict__ us, int32_t* __restrict__ is) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_items) { us[i] = user; is[i] = i; }
}

__global__ void reco_mask_kernel(int32_t user, int32_t n_users, int32_t n_items, const int32_t* __restrict__ urow,
                                 const int32_t* __restrict__ ucol, const double* __restrict__ xcount, double* __restrict__ score) {
  const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_items && !(xcount[t] > 0.0)) score[t] = -INFINITY;  // not in ratings.map(_.item).toSet
  if (user >= 0 && user < n_users) {
    const int32_t b = urow[user], e = urow[user + 1];
    // rated items are few: every thread strides over them (runs after the count mask of its own slot; -inf either way)
    for (int32_t p = b + t; p < e; p += gridDim.x * blockDim.x) score[ucol[p]] = -INFINITY;
  }
}

// single-CTA bitonic sort of P (power of two) (score, item) pairs held in global memory; n valid, the rest padded
__global__ void __launch_bounds__(1024) reco_sort_kernel(const double* __restrict__ score, int32_t n, int32_t P,
                                                        double* __restrict__ key, int32_t* __restrict__ id) {
  for (int32_t x = threadIdx.x; x < P; x += blockDim.x) {
    key[x] = x < n ? score[x] : -INFINITY;
    id[x] = x < n ? x : INT_MAX;
  }
  for (int32_t size = 2; size <= P; size <<= 1) {
    for (int32_t stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int32_t t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const double ka = key[lo], kb = key[hi];
        const int32_t ia = id[lo], ib = id[hi];
        const bool swap = up ? before(kb, ib, ka, ia) : before(ka, ia, kb, ib);
        if (swap) { key[lo] = kb; key[hi] = ka; id[lo] = ib; id[hi] = ia; }
      }
    }
  }
}

int32_t predict_async(const mrs_model* m, const mrs_sim* sim, int32_t kind, const int32_t* d_u, const int32_t* d_i, int64_t n, double* d_out) {
  if (kind == MRS_PRED_PERSONALIZED || kind == MRS_PRED_WSD) {
    MRS_REQUIRE(sim, MRS_ERR_INVALID, "the personalized predictor needs a similarity handle (arbitrary closures cannot run on the GPU)");
    return predict_personalized_async(m, sim, d_u, d_i, n, d_out, kind == MRS_PRED_WSD);
  }
  return predict_baseline_async(m, kind, d_u, d_i, n, d_out);
}

// fp64 FMA throughput of the device (the denominator of the kNN similarity rooflines: there is no fp64 figure in
// MEASURED_PEAKS.json): 8 independent DFMA chains per thread, all SMs full
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* __restrict__ out, int iters) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 1.0 + 1e-9 * (threadIdx.x + k);
  const double b = 1.0000001, c = 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}

}  // namespace
}  // namespace mrs

using namespace mrs;

// measured fp64 FMA rate of the engine's device in FMA/s (diagnostics; bench.py uses it as the kNN roofline denominator)
extern "C" int32_t mrs_debug_fp64_fma_per_s(mrs_engine* e, double* out) {
  MRS_REQUIRE(e && out, MRS_ERR_INVALID, "mrs_debug_fp64_fma_per_s: NULL argument");
  use_engine(e);
  double* d = nullptr;
  MRS_CUDA(cudaMalloc((void**)&d, sizeof(double)));
  cudaEvent_t a, b;
  MRS_CUDA(cudaEventCreate(&a));
  MRS_CUDA(cudaEventCreate(&b));
  const int iters = 20000, grid = e->sm_count * 8;
  fp64_peak_kernel<<<grid, 256, 0, e->stream>>>(d, 1000);  // warm-up
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    MRS_CUDA(cudaEventRecord(a, e->stream));
    fp64_peak_kernel<<<grid, 256, 0, e->stream>>>(d, iters);
    MRS_CUDA(cudaEventRecord(b, e->stream));
    MRS_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    MRS_CUDA(cudaEventElapsedTime(&ms, a, b));
    best = ms < best ? ms : best;
  }
  count_launch(4);
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
  *out = (double)grid * 256.0 * 8.0 * (double)iters / ((double)best * 1e-3);
  return MRS_OK;
}

extern "C" int32_t mrs_mae_async(const mrs_model* m, const mrs_sim* sim, int32_t kind, const mrs_ratings* test, void* device_out2) {
  MRS_REQUIRE(m && test && device_out2, MRS_ERR_INVALID, "mrs_mae_async: NULL argument");
  use_engine(m->eng);
  MRS_REQUIRE(m->eng == test->eng, MRS_ERR_INVALID, "mrs_mae: model and test set live on different engines");
  if (kind == MRS_PRED_PERSONALIZED) {
    MRS_REQUIRE(sim, MRS_ERR_INVALID, "mrs_mae: the personalized predictor needs a similarity handle");
    return mae_personalized_async(m, sim, test, (double*)device_out2);
  }
  return mae_baseline_async(m, kind, test, (double*)device_out2);
}

// MeanAbsoluteErrorSpark(baselinePredictorSpark(train), test) / MAE(computePrediction(train), test) as ONE closure
// (distributed/DistributedBaseline.scala:45-47, predict/Baseline.scala:66-69): three kernels -- user sums, item pass, and a
// test pass that finishes the fit in its prologue -- instead of four.  Falls back to mrs_fit_async + mrs_mae_async for
// rating sets the tiled kernels do not take (fp64-valued ratings, an empty test set, item averages switched on).
extern "C" int32_t mrs_fit_mae_async(mrs_engine* e, const mrs_ratings* train, mrs_model** inout, const mrs_ratings* test, void* device_out2) {
  MRS_REQUIRE(e && train && inout && test && device_out2, MRS_ERR_INVALID, "mrs_fit_mae_async: NULL argument");
  MRS_REQUIRE(train->eng == e && test->eng == e, MRS_ERR_INVALID, "mrs_fit_mae_async: the rating sets live on another engine");
  use_engine(e);
  const bool tiled = train->value_kind == kValueCode && test->value_kind == kValueCode && test->n > 0 && train->n > 0 && *inout &&
                     !(*inout)->want_item_avg;
  if (!tiled) {
    MRS_TRY(fit_local(e, train, inout, true));
    return mae_baseline_async(*inout, MRS_PRED_BASELINE, test, (double*)device_out2);
  }
  MRS_TRY(fit_local(e, train, inout, false, nullptr, true));
  MRS_TRY(launch_mae_tiled_baseline(*inout, test, (double*)device_out2, nullptr, true));
  (*inout)->finished = true;   // the test pass has written the model's arrays
  (*inout)->host_valid = false;
  return MRS_OK;
}

extern "C" int32_t mrs_mae(const mrs_model* m, const mrs_sim* sim, int32_t kind, const mrs_ratings* test, double* mae_out) {
  MRS_REQUIRE(m && test && mae_out, MRS_ERR_INVALID, "mrs_mae: NULL argument");
  mrs_engine* e = m->eng;
  use_engine(e);
  double* d_out = nullptr;
  MRS_TRY(dev_alloc(&d_out, 2));
  int32_t s = mrs_mae_async(m, sim, kind, test, d_out);
  if (s == MRS_OK) {
    cudaError_t ce = cudaMemcpyAsync(e->h_pinned, d_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) { set_error("mrs_mae: %s", cudaGetErrorString(ce)); s = MRS_ERR_CUDA; }
  }
  dev_free(d_out);
  if (s != MRS_OK) return s;
  *mae_out = e->h_pinned[0] / e->h_pinned[1];  // P:85 res._1/res._2 (0.0/0 -> NaN for an empty set)
  return MRS_OK;
}

extern "C" int32_t mrs_predict(const mrs_model* m, const mrs_sim* sim, int32_t kind, const int32_t* users, const int32_t* items,
                               int64_t n, double* out) {
  MRS_REQUIRE(m && (n == 0 || (users && items && out)), MRS_ERR_INVALID, "mrs_predict: NULL argument");
  MRS_REQUIRE(n >= 0, MRS_ERR_INVALID, "mrs_predict: negative n");
  if (n == 0) return MRS_OK;
  mrs_engine* e = m->eng;
  use_engine(e);
  int32_t *d_u = nullptr, *d_i = nullptr;
  double* d_o = nullptr;
  int32_t s = dev_alloc(&d_u, (size_t)n);
  if (s == MRS_OK) s = dev_alloc(&d_i, (size_t)n);
  if (s == MRS_OK) s = dev_alloc(&d_o, (size_t)n);
  if (s == MRS_OK) {
    cudaError_t ce = cudaMemcpyAsync(d_u, users, sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_i, items, sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->stream);
    if (ce != cudaSuccess) { set_error("mrs_predict: %s", cudaGetErrorString(ce)); s = MRS_ERR_CUDA; }
  }
  if (s == MRS_OK) s = predict_async(m, sim, kind, d_u, d_i, n, d_o);
  if (s == MRS_OK) {
    cudaError_t ce = cudaMemcpyAsync(out, d_o, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) { set_error("mrs_predict: %s", cudaGetErrorString(ce)); s = MRS_ERR_CUDA; }
  }
  dev_free(d_u); dev_free(d_i); dev_free(d_o);
  return s;
}

extern "C" int32_t mrs_recommend(const mrs_model* m, const mrs_sim* sim, int32_t kind, int32_t user, int32_t n, int32_t* items_out,
                                 double* scores_out, int32_t* n_out) {
  MRS_REQUIRE(m && n_out && (n <= 0 || (items_out && scores_out)), MRS_ERR_INVALID, "mrs_recommend: NULL argument");
  MRS_REQUIRE(m->finished, MRS_ERR_INVALID, "mrs_recommend: model not finished");
  *n_out = 0;
  if (n <= 0) return MRS_OK;
  mrs_engine* e = m->eng;
  const mrs_ratings* R = m->train;
  const int32_t NI = m->n_items;
  int32_t P = 2;
  while (P < NI) P <<= 1;
  MRS_REQUIRE(P <= (1 << 20), MRS_ERR_UNSUPPORTED, "mrs_recommend: item dimension %d too large for the single-block sort", NI);
  use_engine(e);
  cudaStream_t st = e->stream;
  int32_t *d_u = nullptr, *d_i = nullptr, *d_id = nullptr;
  double *d_score = nullptr, *d_key = nullptr;
  int32_t s = dev_alloc(&d_u, (size_t)NI);
  if (s == MRS_OK) s = dev_alloc(&d_i, (size_t)NI);
  if (s == MRS_OK) s = dev_alloc(&d_score, (size_t)NI);
  if (s == MRS_OK) s = dev_alloc(&d_key, (size_t)P);
  if (s == MRS_OK) s = dev_alloc(&d_id, (size_t)P);
  if (s == MRS_OK) {
    reco_pairs_kernel<<<(NI + 255) / 256, 256, 0, st>>>(user, NI, d_u, d_i);
    count_launch();
    s = predict_async(m, sim, kind, d_u, d_i, NI, d_score);
  }
  if (s == MRS_OK) {
    reco_mask_kernel<<<(NI + 255) / 256, 256, 0, st>>>(user, R->n_users, NI, R->urow, R->ucol, m->xbuf + (size_t)NI, d_score);
    reco_sort_kernel<<<1, 1024, 0, st>>>(d_score, NI, P, d_key, d_id);
    count_launch(2);
    const int32_t w = n < NI ? n : NI;
    std::vector<double> hk((size_t)w);
    std::vector<int32_t> hi((size_t)w);
    cudaError_t ce = cudaMemcpyAsync(hk.data(), d_key, sizeof(double) * w, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(hi.data(), d_id, sizeof(int32_t) * w, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) { set_error("mrs_recommend: %s", cudaGetErrorString(ce)); s = MRS_ERR_CUDA; }
    if (s == MRS_OK) {
      int32_t c = 0;
      for (int32_t j = 0; j < w; ++j) {
        if (hk[j] == -INFINITY) break;  // masked slots sort last
        items_out[c] = hi[j];
        scores_out[c] = hk[j];
        ++c;
      }
      *n_out = c;
    }
  }
  dev_free(d_u); dev_free(d_i); dev_free(d_score); dev_free(d_key); dev_free(d_id);
  return s;
}
