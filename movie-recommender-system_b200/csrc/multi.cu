// multi.cu -- several GPUs driven from ONE host process (what a JVM can do: one JNI library, one thread).
//
// distributed/DistributedBaseline.scala:30-47 runs the baseline pass over the partitions of ONE SparkSession; the
// drop-in for a box of B200s is one process that owns all devices.  mrs_multi_* shards the users over the devices
// (contiguous id ranges balanced by rating count, the same rule as sharded.partition_users), keeps one engine, one
// rating-set pair and one model per device, and runs
//     fit_local (every device) -> exchange of the per-item buffer -> fit_finish -> MAE over the device's test pairs
//     -> 16-byte exchange
// with every launch asynchronous on the device's own stream, so the devices work concurrently although one host
// thread issues everything.  The exchange is the same NVLink peer-memory kernel as in the one-process-per-GPU mode
// (exchange.cu); inside one process the peers' buffers are plain pointers once cudaDeviceEnablePeerAccess has been
// called, no IPC handles are needed.
#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

struct mrs_multi {
  int32_t n = 0;
  std::vector<mrs_engine*> eng;
  std::vector<mrs_ratings*> train, test;
  std::vector<mrs_model*> model;
  std::vector<mrs_exchange*> xch;
  std::vector<double*> d_out2;     // per device: {sum |err|, n}
  std::vector<int32_t> bounds;     // device d owns user ids [bounds[d], bounds[d+1])
  int32_t n_users_dim = 0, n_items_dim = 0;
  int64_t xlen = 0;                // exchanged prefix of the model buffers: 2*I + 2
};

using namespace mrs;

extern "C" void mrs_multi_destroy(mrs_multi* m) {
  if (!m) return;
  for (int d = 0; d < m->n; ++d) {
    if ((size_t)d < m->eng.size() && m->eng[d]) { cudaSetDevice(m->eng[d]->device); cudaStreamSynchronize(m->eng[d]->stream); }
  }
  for (auto* x : m->xch) mrs_exchange_destroy(x);
  for (size_t d = 0; d < m->d_out2.size(); ++d)
    if (m->d_out2[d]) { cudaSetDevice(m->eng[d]->device); cudaFree(m->d_out2[d]); }
  for (auto* p : m->model) mrs_model_destroy(p);
  for (auto* p : m->test) mrs_ratings_destroy(p);
  for (auto* p : m->train) mrs_ratings_destroy(p);
  for (auto* p : m->eng) mrs_engine_destroy(p);
  delete m;
}

extern "C" int32_t mrs_multi_create(const int32_t* device_ids, int32_t n_devices, mrs_multi** out) {
  MRS_REQUIRE(device_ids && out && n_devices >= 1 && n_devices <= 16, MRS_ERR_INVALID, "mrs_multi_create: bad argument (1 <= devices <= 16)");
  mrs_multi* m = new mrs_multi();
  m->n = n_devices;
  for (int d = 0; d < n_devices; ++d) {
    for (int k = 0; k < d; ++k)
      if (device_ids[k] == device_ids[d]) { mrs_multi_destroy(m); set_error("mrs_multi_create: device %d listed twice", device_ids[d]); return MRS_ERR_INVALID; }
    mrs_engine* e = nullptr;
    int32_t s = mrs_engine_create(device_ids[d], nullptr, &e);
    if (s != MRS_OK) { mrs_multi_destroy(m); return s; }
    m->eng.push_back(e);
  }
  // every device maps every other one (the exchange kernel stores flags into and loads partial sums from its peers)
  for (int a = 0; a < n_devices; ++a)
    for (int b = 0; b < n_devices; ++b) {
      if (a == b) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, device_ids[a], device_ids[b]);
      if (!can) { mrs_multi_destroy(m); set_error("mrs_multi_create: device %d cannot access device %d (no NVLink / P2P path)", device_ids[a], device_ids[b]); return MRS_ERR_UNSUPPORTED; }
      cudaSetDevice(device_ids[a]);
      cudaError_t ce = cudaDeviceEnablePeerAccess(device_ids[b], 0);
      if (ce == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (ce != cudaSuccess) { mrs_multi_destroy(m); set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", device_ids[a], device_ids[b], cudaGetErrorString(ce)); return MRS_ERR_CUDA; }
    }
  *out = m;
  return MRS_OK;
}

// contiguous user ranges whose rating counts are as equal as a contiguous split allows (== sharded.partition_users)
static std::vector<int32_t> partition_users(const std::vector<int64_t>& counts, int32_t world) {
  std::vector<int64_t> csum(counts.size());
  std::partial_sum(counts.begin(), counts.end(), csum.begin());
  const int64_t total = csum.empty() ? 0 : csum.back();
  std::vector<int32_t> b{0};
  for (int r = 1; r < world; ++r) {
    const int64_t target = total * r / world;
    int32_t cut = total ? (int32_t)(std::lower_bound(csum.begin(), csum.end(), target) - csum.begin()) + 1 : 0;
    cut = std::min<int32_t>(std::max(cut, b.back()), (int32_t)counts.size());
    b.push_back(cut);
  }
  b.push_back((int32_t)counts.size());
  return b;
}

extern "C" int32_t mrs_multi_load(mrs_multi* m, const int32_t* tr_users, const int32_t* tr_items, const double* tr_ratings, int64_t n_train,
                                  const int32_t* te_users, const int32_t* te_items, const double* te_ratings, int64_t n_test) {
  MRS_REQUIRE(m && (n_train == 0 || (tr_users && tr_items && tr_ratings)) && (n_test == 0 || (te_users && te_items && te_ratings)), MRS_ERR_INVALID,
              "mrs_multi_load: NULL argument");
  MRS_REQUIRE(m->train.empty(), MRS_ERR_INVALID, "mrs_multi_load: rating sets are already loaded");
  int32_t umax = -1, imax = -1;
  for (int64_t k = 0; k < n_train; ++k) {
    MRS_REQUIRE(tr_users[k] >= 0 && tr_items[k] >= 0, MRS_ERR_INVALID, "mrs_multi_load: negative id");
    umax = std::max(umax, tr_users[k]); imax = std::max(imax, tr_items[k]);
  }
  for (int64_t k = 0; k < n_test; ++k) {
    MRS_REQUIRE(te_users[k] >= 0 && te_items[k] >= 0, MRS_ERR_INVALID, "mrs_multi_load: negative id");
    umax = std::max(umax, te_users[k]); imax = std::max(imax, te_items[k]);
  }
  m->n_users_dim = umax + 1; m->n_items_dim = imax + 1;
  if (m->n_users_dim < 1) m->n_users_dim = 1;
  if (m->n_items_dim < 1) m->n_items_dim = 1;
  std::vector<int64_t> counts((size_t)m->n_users_dim, 0);
  for (int64_t k = 0; k < n_train; ++k) counts[(size_t)tr_users[k]]++;
  m->bounds = partition_users(counts, m->n);
  std::vector<int32_t> owner((size_t)m->n_users_dim);
  for (int d = 0; d < m->n; ++d)
    for (int32_t u = m->bounds[d]; u < m->bounds[d + 1]; ++u) owner[(size_t)u] = d;
  // shard both sets by the owner of the user (file order is kept inside a shard)
  for (int pass = 0; pass < 2; ++pass) {
    const int32_t* U = pass ? te_users : tr_users;
    const int32_t* I = pass ? te_items : tr_items;
    const double* Rr = pass ? te_ratings : tr_ratings;
    const int64_t N = pass ? n_test : n_train;
    std::vector<std::vector<int32_t>> su((size_t)m->n), si((size_t)m->n);
    std::vector<std::vector<double>> sr((size_t)m->n);
    for (int64_t k = 0; k < N; ++k) {
      const int d = owner[(size_t)U[k]];
      su[(size_t)d].push_back(U[k]); si[(size_t)d].push_back(I[k]); sr[(size_t)d].push_back(Rr[k]);
    }
    for (int d = 0; d < m->n; ++d) {
      mrs_ratings* R = nullptr;
      MRS_TRY(mrs_ratings_from_coo(m->eng[d], su[(size_t)d].data(), si[(size_t)d].data(), sr[(size_t)d].data(), (int64_t)su[(size_t)d].size(),
                                   m->n_users_dim, m->n_items_dim, &R));
      (pass ? m->test : m->train).push_back(R);
    }
  }
  // models (first local fit allocates them) and the exchange objects
  m->xlen = 2 * (int64_t)m->n_items_dim + 2;
  m->model.assign((size_t)m->n, nullptr);
  m->d_out2.assign((size_t)m->n, nullptr);
  for (int d = 0; d < m->n; ++d) {
    MRS_TRY(mrs_fit_local(m->eng[d], m->train[d], &m->model[d]));
    MRS_TRY(mrs_model_set_item_averages(m->model[d], 0));   // the baseline predictor never forms per-item rating averages (P:362-391)
    use_engine(m->eng[d]);
    MRS_CUDA(cudaMalloc((void**)&m->d_out2[d], 2 * sizeof(double)));
    MRS_CUDA(cudaMemset(m->d_out2[d], 0, 2 * sizeof(double)));
  }
  if (m->n > 1) {
    m->xch.assign((size_t)m->n, nullptr);
    unsigned char handle[64];
    for (int d = 0; d < m->n; ++d) MRS_TRY(mrs_exchange_create(m->eng[d], m->xlen, d, m->n, handle, &m->xch[d]));
    MRS_TRY(mrs_exchange_connect_local(m->xch.data(), m->n));
  }
  for (int d = 0; d < m->n; ++d) MRS_TRY(mrs_engine_sync(m->eng[d]));
  return MRS_OK;
}

// MeanAbsoluteErrorSpark(baselinePredictorSpark(train), test) over all devices (distributed/DistributedBaseline.scala:45-47)
extern "C" int32_t mrs_multi_baseline_mae(mrs_multi* m, double* mae_out) {
  MRS_REQUIRE(m && mae_out && !m->train.empty(), MRS_ERR_INVALID, "mrs_multi_baseline_mae: no rating sets loaded");
  for (int d = 0; d < m->n; ++d) MRS_TRY(mrs_fit_local(m->eng[d], m->train[d], &m->model[d]));
  for (int d = 0; d < m->n && m->n > 1; ++d) {
    void* xbuf = nullptr; int64_t xn = 0;
    MRS_TRY(mrs_model_exchange_buffer(m->model[d], &xbuf, &xn));
    MRS_TRY(mrs_exchange_allreduce_async(m->xch[d], xbuf, m->xlen));                 // P:267-268 / P:247
  }
  for (int d = 0; d < m->n; ++d) {
    MRS_TRY(mrs_fit_finish(m->model[d]));
    MRS_TRY(mrs_mae_async(m->model[d], nullptr, MRS_PRED_BASELINE, m->test[d], m->d_out2[d]));
  }
  for (int d = 0; d < m->n && m->n > 1; ++d) MRS_TRY(mrs_exchange_allreduce_async(m->xch[d], m->d_out2[d], 2));
  double r[2] = {0.0, 0.0};
  for (int d = 0; d < m->n; ++d) MRS_TRY(mrs_engine_sync(m->eng[d]));
  use_engine(m->eng[0]);
  MRS_CUDA(cudaMemcpy(r, m->d_out2[0], sizeof(r), cudaMemcpyDeviceToHost));
  for (int d = 0; d < m->n && m->n > 1; ++d) {
    int32_t timed_out = 0;
    MRS_TRY(mrs_exchange_status(m->xch[d], &timed_out));
    MRS_REQUIRE(!timed_out, MRS_ERR_CUDA, "mrs_multi_baseline_mae: the exchange on device slot %d timed out", d);
  }
  *mae_out = r[0] / r[1];   // P:85: an empty test set gives NaN like the reference's 0.0/0
  return MRS_OK;
}

// the model of one device for queries (borrowed: destroyed with the multi handle).  After a pass the per-item vectors and
// the global average are the same on every device; per-user averages are known on the device that owns the user
// (mrs_multi_owner).
extern "C" int32_t mrs_multi_model(mrs_multi* m, int32_t slot, mrs_model** out) {
  MRS_REQUIRE(m && out && slot >= 0 && slot < m->n && !m->model.empty(), MRS_ERR_INVALID, "mrs_multi_model: bad argument");
  *out = m->model[(size_t)slot];
  return MRS_OK;
}

extern "C" int32_t mrs_multi_owner(const mrs_multi* m, int32_t user, int32_t* slot_out) {
  MRS_REQUIRE(m && slot_out && !m->bounds.empty(), MRS_ERR_INVALID, "mrs_multi_owner: bad argument");
  int32_t d = (int32_t)(std::upper_bound(m->bounds.begin(), m->bounds.end(), user) - m->bounds.begin()) - 1;
  *slot_out = std::min(std::max(d, 0), m->n - 1);
  return MRS_OK;
}
