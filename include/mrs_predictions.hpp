// mrs_predictions.hpp -- C++17 host-side mirror of the reference's `package object shared.predictions`
// (src/main/scala/shared/predictions.scala, "P:") over the C ABI of libmrs_b200.so.
//
// The reference is JVM code and no JVM exists in this image, so this header is the compiled-language form of the
// facade (the Scala/JNI form is in INTEGRATION.md, the Python form in movie-recommender-system_b200/predictions.py):
// same function names, argument meaning and fallbacks.  The reference returns closures; here every factory returns a
// TAGGED function object carrying an engine handle and a kind, so MAE / recommendations issue ONE fused native call and
// operator()(u, i) answers single probes.  There is no CPU fallback: anything that is not a tagged object does not
// type-check, and every native failure is thrown as std::runtime_error(mrs_last_error()).
//
//   g++ -std=c++17 -Iinclude app.cpp -Lmovie-recommender-system_b200 -lmrs_b200 -Wl,-rpath,...
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "mrs_b200.h"

namespace shared {
namespace predictions {

struct Rating {  // P:9
  int user;
  int item;
  double rating;
};

inline void check(int32_t status) {
  if (status != MRS_OK) throw std::runtime_error(mrs_last_error());
}

// ------------------------------------------------------------------ utilities (P:11-61)
template <typename F>
std::pair<double, double> timingInMs(F f) {  // P:11
  const auto start = std::chrono::steady_clock::now();
  const double out = f();
  const auto end = std::chrono::steady_clock::now();
  return {out, std::chrono::duration<double, std::milli>(end - start).count()};
}
inline double mean(const std::vector<double>& s) {  // P:18
  if (s.empty()) return 0.0;
  double acc = s[0];
  for (size_t j = 1; j < s.size(); ++j) acc = acc + s[j];
  return acc / (double)s.size();
}
inline double std_(const std::vector<double>& s) {  // P:19-25 (`std` clashes with the namespace)
  if (s.empty()) return 0.0;
  const double m = mean(s);
  double acc = 0.0;
  for (double x : s) acc += std::pow(m - x, 2);
  return std::sqrt(acc / (double)s.size());
}
inline double scale(double x, double y) {  // P:57-61
  if (x > y) return 5 - y;
  else if (x < y) return y - 1;
  else return 1;
}

// ------------------------------------------------------------------ engine and rating sets
class Engine {
 public:
  explicit Engine(int device = 0) { check(mrs_engine_create(device, nullptr, &h_)); }
  ~Engine() { mrs_engine_destroy(h_); }
  Engine(const Engine&) = delete;
  Engine& operator=(const Engine&) = delete;
  mrs_engine* handle() const { return h_; }
  static Engine& instance() {
    static Engine e(0);
    return e;
  }

 private:
  mrs_engine* h_ = nullptr;
};

// A device-resident Seq[Rating] / RDD[Rating] with its (re-fitted on demand) baseline model.
class RatingSet {
 public:
  explicit RatingSet(const std::vector<Rating>& ratings, Engine& e = Engine::instance()) : eng_(&e) {
    std::vector<int32_t> u(ratings.size()), i(ratings.size());
    std::vector<double> r(ratings.size());
    for (size_t j = 0; j < ratings.size(); ++j) { u[j] = ratings[j].user; i[j] = ratings[j].item; r[j] = ratings[j].rating; }
    check(mrs_ratings_from_coo(e.handle(), u.data(), i.data(), r.data(), (int64_t)ratings.size(), 0, 0, &h_));
  }
  RatingSet(const int32_t* users, const int32_t* items, const double* ratings, int64_t n, Engine& e = Engine::instance()) : eng_(&e) {
    check(mrs_ratings_from_coo(e.handle(), users, items, ratings, n, 0, 0, &h_));
  }
  RatingSet(const std::string& path, const std::string& sep, Engine& e = Engine::instance()) : eng_(&e) {  // load, P:35-49
    check(mrs_ratings_from_file(e.handle(), path.c_str(), sep.c_str(), &h_));
  }
  ~RatingSet() {
    for (auto& kv : sims_) mrs_sim_destroy(kv.second);
    mrs_model_destroy(model_);
    mrs_ratings_destroy(h_);
  }
  RatingSet(const RatingSet&) = delete;
  RatingSet& operator=(const RatingSet&) = delete;

  int64_t length() const {
    int64_t n = 0;
    check(mrs_ratings_info(h_, &n, nullptr, nullptr, nullptr));
    return n;
  }
  mrs_ratings* handle() const { return h_; }
  // the eager part of every computeX(ratings): run the fit kernels (again) on the same buffers
  mrs_model* fit() const {
    check(mrs_fit_async(eng_->handle(), h_, &model_));
    return model_;
  }
  mrs_model* model() const { return model_ ? model_ : fit(); }
  // similarities of one kind are computed once per rating set; k only selects a prefix of the sorted lists (P:626)
  mrs_sim* similarity(int kind, int k) const {
    auto it = sims_.find(kind);
    if (it != sims_.end()) {
      check(mrs_sim_set_k(it->second, k));
      return it->second;
    }
    mrs_sim* s = nullptr;
    check(mrs_fit_similarity(model(), kind, k, &s));
    sims_[kind] = s;
    return s;
  }

 private:
  Engine* eng_;
  mrs_ratings* h_ = nullptr;
  mutable mrs_model* model_ = nullptr;
  mutable std::map<int, mrs_sim*> sims_;
};

inline RatingSet load(const std::string& path, const std::string& sep) { return RatingSet(path, sep); }  // P:35 (no SparkSession)

// ------------------------------------------------------------------ tagged function objects
struct Similarity {  // (Int, Int) => Double
  const RatingSet* train = nullptr;
  int kind = MRS_SIM_UNIFORM;
  int k = 0;
  mrs_sim* handle = nullptr;
  double operator()(int u, int v) const {
    if (kind == MRS_SIM_UNIFORM) return 1.0;  // P:400
    check(mrs_sim_set_k(handle, k));
    double out = 0.0;
    check(mrs_similarity(handle, u, v, &out));
    return out;
  }
};

struct Predictor {  // (Int, Int) => Double
  const RatingSet* train = nullptr;
  int kind = MRS_PRED_BASELINE;
  Similarity sim;
  double operator()(int u, int i) const {
    if (sim.handle) check(mrs_sim_set_k(sim.handle, sim.k));
    const int32_t uu = u, ii = i;
    double out = 0.0;
    check(mrs_predict(train->model(), sim.handle, kind, &uu, &ii, 1, &out));
    return out;
  }
};

struct WeightedSumDeviation {  // (Int, Int) => Double
  const RatingSet* train = nullptr;
  Similarity sim;
  double operator()(int u, int i) const {
    check(mrs_sim_set_k(sim.handle, sim.k));
    const int32_t uu = u, ii = i;
    double out = 0.0;
    check(mrs_predict(train->model(), sim.handle, MRS_PRED_WSD, &uu, &ii, 1, &out));
    return out;
  }
};

// ------------------------------------------------------------------ baseline family (P:69-237) and Spark twins (P:246-391)
inline double MAE(const Predictor& predict, const RatingSet& data) {  // P:69-86: one fused native call
  if (predict.sim.handle) check(mrs_sim_set_k(predict.sim.handle, predict.sim.k));
  double out = 0.0;
  check(mrs_mae(predict.train->model(), predict.sim.handle, predict.kind, data.handle(), &out));
  return out;
}
inline double average(const RatingSet& ratings) {  // P:94
  double out = 0.0;
  check(mrs_model_scalar(ratings.fit(), MRS_GLOBAL_AVG, &out));
  return out;
}
inline Predictor computeAvgRating(const RatingSet& r) { r.fit(); return {&r, MRS_PRED_GLOBAL, {}}; }    // P:101
inline Predictor computeUserAvg(const RatingSet& r) { r.fit(); return {&r, MRS_PRED_USER, {}}; }        // P:120
inline Predictor computeItemAvg(const RatingSet& r) { r.fit(); return {&r, MRS_PRED_ITEM, {}}; }        // P:141
inline Predictor computeItemAvgDev(const RatingSet& r) { r.fit(); return {&r, MRS_PRED_ITEMDEV, {}}; }  // P:193
inline Predictor computePrediction(const RatingSet& r) { r.fit(); return {&r, MRS_PRED_BASELINE, {}}; } // P:205

inline std::map<int, double> vector_of(const RatingSet& r, int kind) {
  mrs_model* m = r.fit();
  int64_t n = 0;
  check(mrs_model_vector(m, kind, nullptr, nullptr, 0, &n));
  std::vector<double> vals((size_t)n);
  std::vector<int32_t> cnt((size_t)n);
  check(mrs_model_vector(m, kind, vals.data(), cnt.data(), n, &n));
  std::map<int, double> out;
  for (int64_t j = 0; j < n; ++j)
    if (cnt[(size_t)j] > 0) out[(int)j] = vals[(size_t)j];
  return out;
}
inline std::map<int, double> usersAvg(const RatingSet& r) { return vector_of(r, MRS_USER_AVG); }         // P:113
inline std::map<int, double> itemsAvg(const RatingSet& r) { return vector_of(r, MRS_ITEM_AVG); }         // P:134
inline std::map<int, double> itemsAvgDev(const RatingSet& r) { return vector_of(r, MRS_ITEM_AVG_DEV); }  // P:176

inline double MeanAbsoluteErrorSpark(const Predictor& p, const RatingSet& real) { return MAE(p, real); }  // P:256
inline double getGlobalAvg(const RatingSet& r) { return average(r); }                                     // P:265
inline std::map<int, double> getUsersAvg(const RatingSet& r) { return usersAvg(r); }                      // P:274
inline Predictor usersAvgSpark(const RatingSet& r) { return computeUserAvg(r); }                          // P:281
inline std::map<int, double> getItemsAvg(const RatingSet& r) { return itemsAvg(r); }                      // P:295
inline Predictor itemsAvgSpark(const RatingSet& r) { return computeItemAvg(r); }                          // P:302
inline std::map<int, double> getItemsAvgDev(const RatingSet& r) { return itemsAvgDev(r); }                // P:336
inline Predictor itemsAvgDevSpark(const RatingSet& r) { return computeItemAvgDev(r); }                    // P:350
inline Predictor baselinePredictorSpark(const RatingSet& r) { return computePrediction(r); }              // P:362

// ------------------------------------------------------------------ personalized / kNN (P:400-674)
inline Similarity similarityOne() { return {}; }                                                         // P:400
inline Similarity adjustedCosineSimilarityFunction(const RatingSet& r) {                                  // P:407
  return {&r, MRS_SIM_COSINE, 0, r.similarity(MRS_SIM_COSINE, 0)};
}
inline Similarity jaccardCoefficient(const RatingSet& r) {                                                // P:440
  return {&r, MRS_SIM_JACCARD, 0, r.similarity(MRS_SIM_JACCARD, 0)};
}
inline Similarity getSimilarity(const RatingSet& r, int k, const Similarity& s) {                         // P:626
  return {&r, s.kind, k, r.similarity(s.kind, k)};
}
inline WeightedSumDeviation weightedSumDeviation(const RatingSet& r, const Similarity& s) {               // P:489
  return {&r, {&r, s.kind, s.k, r.similarity(s.kind, s.k)}};
}
inline Predictor predictor(const RatingSet& r, const WeightedSumDeviation& wsd) {                         // P:557
  if (wsd.train == &r) return {&r, MRS_PRED_PERSONALIZED, wsd.sim};
  return {&r, MRS_PRED_PERSONALIZED, {&r, wsd.sim.kind, wsd.sim.k, r.similarity(wsd.sim.kind, wsd.sim.k)}};
}
inline std::vector<std::pair<int, double>> getNeighbors(const RatingSet& r, int k, const Similarity& s, int u) {  // P:596 applied to u
  mrs_sim* h = r.similarity(s.kind, k);
  std::vector<int32_t> ids((size_t)std::max(k, 1));
  std::vector<double> sims((size_t)std::max(k, 1));
  int32_t n = 0;
  check(mrs_neighbors(h, u, k, ids.data(), sims.data(), (int32_t)ids.size(), &n));
  std::vector<std::pair<int, double>> out;
  for (int32_t j = 0; j < n; ++j) out.emplace_back(ids[(size_t)j], sims[(size_t)j]);
  return out;
}
inline std::vector<std::pair<int, double>> recommendations(const RatingSet& r, const Predictor& p, int user, int n) {  // P:651 applied to (user, n)
  (void)r;
  if (p.sim.handle) check(mrs_sim_set_k(p.sim.handle, p.sim.k));
  std::vector<int32_t> items((size_t)std::max(n, 1));
  std::vector<double> scores((size_t)std::max(n, 1));
  int32_t w = 0;
  check(mrs_recommend(p.train->model(), p.sim.handle, p.kind, user, n, items.data(), scores.data(), &w));
  std::vector<std::pair<int, double>> out;
  for (int32_t j = 0; j < w; ++j) out.emplace_back(items[(size_t)j], scores[(size_t)j]);
  return out;
}

}  // namespace predictions
}  // namespace shared
