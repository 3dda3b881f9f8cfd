// mirror_check.cpp -- drives include/mrs_predictions.hpp the way the reference's mains drive shared.predictions
// (predict/Baseline.scala:40-67, predict/Personalized.scala:40-67, predict/kNN.scala:40-44,
// recommend/Recommender.scala:82-88) and prints one JSON object; tests/test_gpu_cpp_mirror.py compares it with the oracle.
#include <cstdio>
#include <cstdlib>

#include "mrs_predictions.hpp"

using namespace shared::predictions;

static void print_pairs(const char* key, const std::vector<std::pair<int, double>>& v, const char* tail) {
  std::printf("\"%s\": [", key);
  for (size_t j = 0; j < v.size(); ++j) std::printf("%s[%d, %.17g]", j ? ", " : "", v[j].first, v[j].second);
  std::printf("]%s", tail);
}

int main(int argc, char** argv) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: mirror_check train test sep k user\n");
    return 2;
  }
  try {
    RatingSet train = load(argv[1], argv[3]);
    RatingSet test = load(argv[2], argv[3]);
    const int k = std::atoi(argv[4]), user = std::atoi(argv[5]);
    std::printf("{\"n_train\": %lld, \"n_test\": %lld, ", (long long)train.length(), (long long)test.length());
    std::printf("\"global_avg\": %.17g, ", average(train));
    std::printf("\"mae_global\": %.17g, ", MAE(computeAvgRating(train), test));
    std::printf("\"mae_user\": %.17g, ", MAE(computeUserAvg(train), test));
    std::printf("\"mae_item\": %.17g, ", MAE(computeItemAvg(train), test));
    std::printf("\"mae_itemdev\": %.17g, ", MAE(computeItemAvgDev(train), test));
    std::printf("\"mae_baseline\": %.17g, ", MAE(computePrediction(train), test));
    std::printf("\"mae_baseline_spark\": %.17g, ", MeanAbsoluteErrorSpark(baselinePredictorSpark(train), test));
    std::printf("\"pred_baseline_1_1\": %.17g, ", computePrediction(train)(1, 1));
    std::printf("\"pred_unknown_user\": %.17g, ", computePrediction(train)(1 << 20, 1));
    const auto ua = usersAvg(train);
    std::printf("\"n_users\": %zu, \"user_avg_first\": [%d, %.17g], ", ua.size(), ua.begin()->first, ua.begin()->second);

    const Similarity cosine = adjustedCosineSimilarityFunction(train);
    const Similarity jaccard = jaccardCoefficient(train);
    std::printf("\"mae_uniform\": %.17g, ", MAE(predictor(train, weightedSumDeviation(train, similarityOne())), test));
    std::printf("\"mae_cosine\": %.17g, ", MAE(predictor(train, weightedSumDeviation(train, cosine)), test));
    std::printf("\"mae_jaccard\": %.17g, ", MAE(predictor(train, weightedSumDeviation(train, jaccard)), test));
    std::printf("\"sim_cosine_1_2\": %.17g, \"sim_jaccard_1_2\": %.17g, ", cosine(1, 2), jaccard(1, 2));
    const Similarity knn = getSimilarity(train, k, cosine);
    const Predictor pk = predictor(train, weightedSumDeviation(train, knn));
    std::printf("\"mae_knn\": %.17g, \"pred_knn_1_1\": %.17g, ", MAE(pk, test), pk(1, 1));
    std::printf("\"wsd_knn_1_1\": %.17g, ", weightedSumDeviation(train, knn)(1, 1));
    print_pairs("neighbors", getNeighbors(train, k, cosine, user), ", ");
    print_pairs("recommendations", recommendations(train, pk, user, 5), ", ");
    std::printf("\"scale\": [%.17g, %.17g, %.17g], \"std\": %.17g}\n", scale(4, 3), scale(2, 3), scale(3, 3), std_({1, 2, 3, 4}));
  } catch (const std::exception& e) {
    std::fprintf(stderr, "mirror_check: %s\n", e.what());
    return 1;
  }
  return 0;
}
