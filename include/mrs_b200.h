/*
 * mrs_b200.h -- C ABI of libmrs_b200.so, the B200 (sm_100a) engine for the rating-prediction hot
 * path of EloDoyard/movie-recommender-system.
 *
 * The reference has no FFI of its own: its boundary is the public surface of the Scala package
 * object `shared.predictions` (src/main/scala/shared/predictions.scala, "P:" below).  Each entry
 * point here names the reference function(s) it replaces; the JNI/Scala binding a maintainer
 * would add is shown in INTEGRATION.md.  Plain pointers and sizes only; no exceptions cross the
 * boundary.  Every function returns 0 (MRS_OK) or a negative mrs_status; mrs_last_error() gives a
 * thread-local message for the last failure on the calling thread.
 *
 * Model of use (fit -> device-resident model -> batched query):
 *   engine  = one CUDA device + one stream (one process per GPU; multi-GPU is sharded by the host
 *             layer with one collective on the exchange buffer, see mrs_fit_local/mrs_fit_finish)
 *   ratings = a device-resident rating set: user-major CSR + item-major CSC (+ sorted COO)
 *   model   = global / per-user / per-item averages, item average deviations      (P:94-237, 246-391)
 *   sim     = user-user similarity (uniform | cosine | jaccard) with optional top-k (P:400-481, 596-649)
 * Ids are the reference's original Int ids (>= 0); tables are direct-indexed by id.
 * Handles are not thread-safe; use one engine per host thread.
 */
#ifndef MRS_B200_H
#define MRS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MRS_API __attribute__((visibility("default")))
#else
#define MRS_API
#endif

typedef struct mrs_engine mrs_engine;
typedef struct mrs_ratings mrs_ratings;
typedef struct mrs_model mrs_model;
typedef struct mrs_sim mrs_sim;
typedef struct mrs_graph mrs_graph;
typedef struct mrs_exchange mrs_exchange;
typedef struct mrs_multi mrs_multi;       /* several devices driven from one process */
typedef struct mrs_upload mrs_upload;     /* host -> device copies of one rating set in flight */

typedef enum {
  MRS_OK = 0,
  MRS_ERR_INVALID = -1,     /* bad argument (NULL handle, negative id, unknown kind ...) */
  MRS_ERR_CUDA = -2,        /* a CUDA runtime call failed (message has the CUDA error string) */
  MRS_ERR_NOMEM = -3,
  MRS_ERR_DUPLICATE = -4,   /* duplicate (user,item) pair in a rating set (P:168 keeps the last; unsupported here) */
  MRS_ERR_IO = -5,          /* file missing / malformed row (the reference throws at P:41) */
  MRS_ERR_UNSUPPORTED = -6  /* shape outside what this build handles; never a silent CPU fallback */
} mrs_status;

/* vectors / scalars of a fitted model */
typedef enum {
  MRS_GLOBAL_AVG = 0,   /* average            P:94,  getGlobalAvg     P:265 */
  MRS_USER_AVG = 1,     /* usersAvg           P:113, getUsersAvg      P:274 */
  MRS_ITEM_AVG = 2,     /* itemsAvg           P:134, getItemsAvg      P:295 */
  MRS_ITEM_AVG_DEV = 3  /* itemsAvgDev        P:176, getItemsAvgDev   P:336 */
} mrs_vec_kind;

/* predictors ((Int,Int) => Double closures of the reference) */
typedef enum {
  MRS_PRED_GLOBAL = 0,       /* computeAvgRating   P:101 */
  MRS_PRED_USER = 1,         /* computeUserAvg     P:120, usersAvgSpark    P:281 */
  MRS_PRED_ITEM = 2,         /* computeItemAvg     P:141, itemsAvgSpark    P:302 */
  MRS_PRED_ITEMDEV = 3,      /* computeItemAvgDev  P:193, itemsAvgDevSpark P:350 */
  MRS_PRED_BASELINE = 4,     /* computePrediction  P:205, baselinePredictorSpark P:362 */
  MRS_PRED_PERSONALIZED = 5, /* predictor(ratings, weightedSumDeviation(ratings, sim))  P:489-586; needs an mrs_sim */
  MRS_PRED_WSD = 6           /* weightedSumDeviation(ratings, sim) itself              P:489-549; needs an mrs_sim; mrs_predict only */
} mrs_pred_kind;

typedef enum {
  MRS_SIM_UNIFORM = 0,  /* similarityOne                      P:400 */
  MRS_SIM_COSINE = 1,   /* adjustedCosineSimilarityFunction   P:407-433 (+ preprocessedRating P:470-481) */
  MRS_SIM_JACCARD = 2   /* jaccardCoefficient                 P:440-464 */
} mrs_sim_kind;

/* ---- library ---- */
MRS_API const char* mrs_last_error(void);
MRS_API const char* mrs_version(void);
/* number of kernel launches issued by this library on this process so far (bench.py's gpu_launches) */
MRS_API int64_t mrs_launch_count(void);

/* ---- engine ---- */
/* cuda_stream: a cudaStream_t to enqueue on (e.g. the caller's current stream), or NULL to create one. */
MRS_API int32_t mrs_engine_create(int32_t device, void* cuda_stream, mrs_engine** out);
MRS_API void mrs_engine_destroy(mrs_engine* e);
MRS_API int32_t mrs_engine_sync(mrs_engine* e);
/* diagnostics: an engine created with MRS_TIMELINE=1 in the environment records, per kernel k of the baseline pass (0 user
 * sums, 1 item pass, 2 item finalize, 3 test pass), the earliest block start out32[2k] and the latest block end out32[2k+1]
 * in %globaltimer nanoseconds; the call synchronises, copies them out and re-arms the slots (tools/timeline.py) */
MRS_API int32_t mrs_debug_timeline(mrs_engine* e, uint64_t* out32);
/* per-CTA stamps of the last baseline pass, 4 x 256 values: item pass "table built" and "done", test pass "table built" and "done" */
MRS_API int32_t mrs_debug_cta_stamps(mrs_engine* e, uint64_t* out1024);
/* per-warp end stamps and (rows << 32 | slices) of the item pass; filled only by a library built with -DMRS_WARP_STAMPS */
MRS_API int32_t mrs_debug_warp_stamps(mrs_engine* e, uint64_t* out16384);
/* diagnostics: measured fp64 FMA rate of the engine's device, FMA per second (denominator of the kNN similarity rooflines) */
MRS_API int32_t mrs_debug_fp64_fma_per_s(mrs_engine* e, double* out);

/* CUDA-graph capture of any sequence of the asynchronous entry points (mrs_fit_async, mrs_fit_local, mrs_fit_finish,
 * mrs_fit_similarity_async, mrs_mae_async) on this engine's stream: the kernels of a pass take tens of microseconds, so
 * replaying one graph instead of launching them one by one removes the launch latency from the step.  Every handle used
 * between begin and end must have been used once before (layouts and buffers are allocated on first use). */
MRS_API int32_t mrs_graph_begin(mrs_engine* e);
MRS_API int32_t mrs_graph_end(mrs_engine* e, mrs_graph** out);
MRS_API int32_t mrs_graph_launch(mrs_graph* g);
MRS_API void mrs_graph_destroy(mrs_graph* g);

/* Per-kernel device timing (diagnostics; what bench.py's roofline uses): between begin and end every kernel the
 * library launches on this engine is bracketed by CUDA events on the engine's stream.  mrs_profile_end syncs and
 * returns one '\n'-separated label per launch in names_out and its duration in milliseconds in ms_out. */
MRS_API int32_t mrs_profile_begin(mrs_engine* e);
MRS_API int32_t mrs_profile_end(mrs_engine* e, char* names_out, int64_t names_cap, float* ms_out, int32_t cap, int32_t* n_out);

/* ---- ratings: replaces `load(...).collect()` / RDD[Rating] materialisation (P:35-49; call sites
 * predict/Baseline.scala:40-42, distributed/DistributedBaseline.scala:41-43) ---- */
/* Inputs are host arrays (Rating.user, Rating.item, Rating.rating); they are copied, never retained.
 * n_users_dim / n_items_dim: minimum table sizes (max id + 1) to allocate, or 0 to derive them from the data;
 * ranks of a sharded run pass the global values so that their exchange buffers line up. */
MRS_API int32_t mrs_ratings_from_coo(mrs_engine* e, const int32_t* users, const int32_t* items, const double* ratings,
                                     int64_t n, int32_t n_users_dim, int32_t n_items_dim, mrs_ratings** out);
/* Staged form of mrs_ratings_from_coo for callers that hand over several sets (train and test): mrs_upload_begin
 * enqueues the host -> device copies on the engine's copy stream and returns at once (ids first, ratings last), so the
 * copies of the second set run while the first one is being built; mrs_ratings_from_upload builds the set (it starts
 * sorting as soon as the ids have arrived) and consumes the upload.  The host arrays must stay valid until
 * mrs_ratings_from_upload (or mrs_upload_destroy) returns; they should be page-locked, otherwise the copies do not overlap.
 * mrs_ratings_from_coo == mrs_upload_begin + mrs_ratings_from_upload. */
MRS_API int32_t mrs_upload_begin(mrs_engine* e, const int32_t* users, const int32_t* items, const double* ratings, int64_t n,
                                 mrs_upload** out);
MRS_API int32_t mrs_ratings_from_upload(mrs_upload* up, int32_t n_users_dim, int32_t n_items_dim, mrs_ratings** out);
/* Compact host form for half-star data: the rating as a 1-byte code = 2 x rating (0 .. 254), i.e. 9 bytes per Rating over
 * PCIe instead of 16 -- the link is what bounds an end-to-end step (the device stores codes anyway).  A JNI caller packs
 * Rating.rating into the byte while it fills its pinned buffers.  mrs_ratings_from_coo_codes == begin + from_upload. */
MRS_API int32_t mrs_upload_begin_codes(mrs_engine* e, const int32_t* users, const int32_t* items, const uint8_t* codes, int64_t n,
                                       mrs_upload** out);
MRS_API int32_t mrs_ratings_from_coo_codes(mrs_engine* e, const int32_t* users, const int32_t* items, const uint8_t* codes, int64_t n,
                                           int32_t n_users_dim, int32_t n_items_dim, mrs_ratings** out);
MRS_API void mrs_upload_destroy(mrs_upload* up);
/* Same parse rules as P:35-49: split on `sep`, trim, keep the row iff column 0 parses as an Int. */
MRS_API int32_t mrs_ratings_from_file(mrs_engine* e, const char* path, const char* sep, mrs_ratings** out);
/* The same on `nbytes` of text in host memory (what a JVM caller holds after reading a file or an HDFS block).  The text is
 * copied to the device and parsed there, one thread per line; only ratings written in another form than
 * [sign]digits[.digits] with at most 15 digits (exponents, hex, NaN ...) send the text through the host parser instead.
 * `sep` is a literal string of at most 8 bytes (the reference passes it to String.split, i.e. as a regex: "\t", ",", "::"
 * mean the same either way). */
MRS_API int32_t mrs_ratings_from_text(mrs_engine* e, const char* text, int64_t nbytes, const char* sep, mrs_ratings** out);
/* value_kind: 0 = every rating is a multiple of 0.5 in [0,127.5] and is stored as a 1-byte code; 1 = fp64 values */
MRS_API int32_t mrs_ratings_info(const mrs_ratings* r, int64_t* n, int32_t* n_users_dim, int32_t* n_items_dim, int32_t* value_kind);
/* bytes of device memory the hot kernels read per pass over this set: [0] user-major payload, [1] item-major payload,
 * [2] sorted-COO payload (the figures bench.py's roofline uses) */
MRS_API int32_t mrs_ratings_bytes(const mrs_ratings* r, int64_t* bytes3);
/* diagnostics: sizes of the lazily built kernel layouts (0 until built):
 * [0] user tiles, [1] units, [2] slices, [3] slots of the tiled item-major layout; [4] item tiles, [5] chunks, [6] slots of the
 * item-tiled test layout; [7] 16-code vectors of the padded user-major code array */
MRS_API int32_t mrs_ratings_layout_info(const mrs_ratings* r, int64_t* out8);
MRS_API void mrs_ratings_destroy(mrs_ratings* r);

/* ---- fit: replaces the eager part of computePrediction (P:205-214) / baselinePredictorSpark (P:362-368)
 * and of computeAvgRating/computeUserAvg/computeItemAvg/computeItemAvgDev ---- */
/* `train` must outlive the model.  mrs_fit == mrs_fit_local + mrs_fit_finish (+ a stream sync). */
MRS_API int32_t mrs_fit(mrs_engine* e, const mrs_ratings* train, mrs_model** out);
/* Asynchronous pieces (enqueue on the engine's stream, no host sync); *out may be an existing model of the same
 * train set, in which case its buffers are reused (this is what a timed loop calls). */
MRS_API int32_t mrs_fit_local(mrs_engine* e, const mrs_ratings* train, mrs_model** inout);
/* single-GPU asynchronous fit: mrs_fit_local and mrs_fit_finish fused (no exchange step, one kernel fewer) */
MRS_API int32_t mrs_fit_async(mrs_engine* e, const mrs_ratings* train, mrs_model** inout);
/* Users-only refit of an existing, fitted model of `train` (asynchronous): the user averages and the global average, which is
 * all that predictor(ratings, weightedSumDeviation(...)) takes from the fit (P:557-586, P:489-549).  The per-item average
 * deviation keeps the values of the last full fit (a model's train set never changes).  This is the fit of the timed kNN
 * closure (predict/kNN.scala:42-45). */
MRS_API int32_t mrs_fit_users_async(mrs_engine* e, const mrs_ratings* train, mrs_model** inout);
/* The exchange buffer written by mrs_fit_local: n_doubles = 3*I+2 fp64 values on the device (I = item table size),
 * [ sum of deviations per item | count per item | sum of all ratings | count | sum of ratings per item ].
 * A sharded run all-reduces (sum) it across ranks between mrs_fit_local and mrs_fit_finish: this is the one
 * collective that replaces the reduceByKey/collect shuffles of P:267-268 and the sum/count of P:247.  The first
 * 2*I+2 values are all the baseline predictor needs: with item averages switched off only that prefix has to travel. */
/* Per-item rating averages (itemsAvg, P:134) are accumulated in the same pass as the deviations by default.  The
 * baseline predictor (P:205 / P:362) does not use them: a caller that only needs that predictor can switch them off
 * for subsequent fits of this model; item-average queries then fail with MRS_ERR_INVALID. */
MRS_API int32_t mrs_model_set_item_averages(mrs_model* m, int32_t enabled);
MRS_API int32_t mrs_model_exchange_buffer(mrs_model* m, void** device_ptr, int64_t* n_doubles);
MRS_API int32_t mrs_fit_finish(mrs_model* m);
MRS_API void mrs_model_destroy(mrs_model* m);

/* ---- the collective of a sharded run as our own kernel over NVLink peer memory (one process per GPU, one box).
 * Each rank creates an exchange object (a symmetric device buffer + flags) and gets a 64-byte CUDA-IPC handle; the
 * host layer all-gathers the handles (any transport) and passes the world x 64 bytes to mrs_exchange_connect.
 * mrs_exchange_allreduce_async then sums `n_doubles` fp64 values in place across the ranks IN RANK ORDER (bit-identical
 * result everywhere) with one kernel: publish, flag the peers, wait, 128-bit peer loads.  Replaces the NCCL all-reduce of
 * the exchange buffer (P:267-268, P:247) and of {sum |err|, n}; can be captured in a CUDA graph.  Every rank must issue
 * the same sequence of calls.  A peer that never arrives makes the kernel give up after ~2 s (mrs_exchange_set_timeout_ms):
 * it then overwrites the caller's buffer with NaN (nothing downstream can pass for a result) and sets an error word that
 * mrs_exchange_status reads; once the host has seen it the handle refuses further exchanges (MRS_ERR_CUDA). */
MRS_API int32_t mrs_exchange_create(mrs_engine* e, int64_t n_doubles, int32_t rank, int32_t world, void* ipc_handle_out64, mrs_exchange** out);
MRS_API int32_t mrs_exchange_connect(mrs_exchange* x, const void* all_handles_world_x_64);
/* all ranks are devices of the calling process (peer access enabled): connect xs[0..world) to each other without IPC */
MRS_API int32_t mrs_exchange_connect_local(mrs_exchange** xs, int32_t world);
MRS_API int32_t mrs_exchange_allreduce_async(mrs_exchange* x, void* device_inout, int64_t n_doubles);
/* Same over the positions device_idx[0..n_idx) of `device_inout` only (int32, on the device, identical on every rank): the
 * other positions neither travel nor change.  The exchange buffer of a model is indexed by item id; with sparse ids most of
 * its slots are zero on every rank (72 % at ml-25m shape), and the slots in use are known once the rating sets are loaded. */
MRS_API int32_t mrs_exchange_allreduce_indexed_async(mrs_exchange* x, void* device_inout, const int32_t* device_idx, int64_t n_idx);
/* Fused compute + collective: the exchange happens INSIDE the pass' own kernels instead of between them.
 * mrs_fit_local_push = mrs_fit_local whose last kernel delivers this rank's per-item partial sums straight into every
 * rank's symmetric receive buffer with NVLink stores and raises a flag (no publish copy, no waiting);
 * mrs_fit_finish_pull = mrs_fit_finish that first waits for every rank's delivery and adds them from its OWN memory in
 * rank order (bit-identical totals everywhere, no remote loads); mrs_mae_push_async = the fused baseline MAE whose last
 * block delivers {sum |err|, n} to every rank, waits for the others and writes the global pair to device_out2.
 * `device_known_items`: the n_known item ids (ascending, int32, on the device, identical on every rank) that occur on
 * SOME rank -- only their slots travel.  The model must have item averages switched off.  Each of the two exchanges of a
 * pass needs its own handle (capacities >= 2*n_known+2 and >= 2 doubles); a missing peer poisons the results with NaN
 * and is reported by mrs_exchange_status like in the unfused form. */
MRS_API int32_t mrs_fit_local_push(mrs_engine* e, const mrs_ratings* train, mrs_model** inout, mrs_exchange* x,
                                   const int32_t* device_known_items, int32_t n_known);
MRS_API int32_t mrs_fit_finish_pull(mrs_model* m, mrs_exchange* x);
MRS_API int32_t mrs_mae_push_async(const mrs_model* m, const mrs_ratings* test, mrs_exchange* x, void* device_out2);
MRS_API int32_t mrs_exchange_status(mrs_exchange* x, int32_t* timed_out);
MRS_API int32_t mrs_exchange_set_timeout_ms(mrs_exchange* x, int64_t milliseconds);
/* diagnostics: %globaltimer (ns) of block 0 in the last exchange: [0] start, [1] published, [2] first barrier passed,
 * [3] slice reduced, [4] second barrier passed (two-shot only), [5] done */
MRS_API int32_t mrs_exchange_stamps(mrs_exchange* x, uint64_t* out8);
MRS_API void mrs_exchange_destroy(mrs_exchange* x);

MRS_API int32_t mrs_model_scalar(const mrs_model* m, int32_t vec_kind, double* out);
/* value for one original id with the reference's fallbacks (unknown user/item -> global average; unknown item
 * deviation -> 0.0).  known_out (optional) = 1 iff the id occurs in the train set. */
MRS_API int32_t mrs_model_lookup(const mrs_model* m, int32_t vec_kind, int32_t id, double* out, int32_t* known_out);
/* whole table, direct-indexed by id (entries of unknown ids hold the fallback); counts_out optional */
MRS_API int32_t mrs_model_vector(const mrs_model* m, int32_t vec_kind, double* vals_out, int32_t* counts_out, int64_t cap, int64_t* n_out);

/* ---- similarity: replaces adjustedCosineSimilarityFunction / jaccardCoefficient / similarityOne and
 * getNeighbors / getSimilarity (P:596-649).  k <= 0: no neighbourhood restriction. ---- */
MRS_API int32_t mrs_fit_similarity(mrs_model* m, int32_t sim_kind, int32_t k, mrs_sim** out);
MRS_API int32_t mrs_fit_similarity_async(mrs_model* m, int32_t sim_kind, int32_t k, mrs_sim** inout);
/* Row-block form (always the list path, any user count): neighbour lists are computed and kept only for the users with
 * original id in [user_lo, user_hi) -- the rows one rank of a sharded run owns (BASELINE config 5; every rank holds the
 * whole train set, so no exchange is needed before the rows).  Only the first min(k, users-1) neighbours of a user are
 * kept (k > 0): mrs_sim_set_k accepts 0 < k' <= k, queries for users outside the range fail (mrs_predict / mrs_mae write
 * NaN for them).  mrs_fit_similarity itself takes this path over all users above 16,384 users. */
MRS_API int32_t mrs_fit_similarity_rows_async(mrs_model* m, int32_t sim_kind, int32_t k, int32_t user_lo, int32_t user_hi, mrs_sim** inout);
/* Order of neighbours whose similarities are EXACTLY equal (typically the block of zero similarities that a large k reaches).
 * The reference's stable sort (P:610) keeps the candidate order of `(allUsers - u).toSeq` (P:608), i.e. the iteration order of
 * a Scala 2.11 immutable.HashSet[Int]; BASELINE's north_star asks for index order.  mode 0 (default): ascending user id;
 * mode 1: that HashSet order (hash-trie walk of improve(id), SURVEY A.6 -- recalled from the 2.11 library source, not
 * verifiable without a JVM).  Applies to similarity handles fitted AFTER the call.  Only exact ties move. */
MRS_API int32_t mrs_model_set_tie_order(mrs_model* m, int32_t mode);
/* change k without recomputing similarities (the sorted lists have the prefix property) */
MRS_API int32_t mrs_sim_set_k(mrs_sim* s, int32_t k);
MRS_API int32_t mrs_similarity(const mrs_sim* s, int32_t u, int32_t v, double* out);
/* first min(k, candidates) neighbours of u, order (similarity desc, user id asc); P:603-616 */
MRS_API int32_t mrs_neighbors(const mrs_sim* s, int32_t u, int32_t k, int32_t* ids_out, double* sims_out, int32_t cap, int32_t* n_out);
/* per-rating values in user-major order (user asc, item asc): which = 0 -> computeNormalizeDeviation (P:155-169),
 * which = 1 -> preprocessedRating (P:470-481).  Any of the output arrays may be NULL; n_out gets the entry count. */
MRS_API int32_t mrs_sim_entry_values(const mrs_sim* s, int32_t which, int32_t* users_out, int32_t* items_out, double* vals_out,
                                     int64_t cap, int64_t* n_out);
MRS_API void mrs_sim_destroy(mrs_sim* s);

/* ---- batched prediction and fused MAE: replace `predict(u,i)` closures and MAE / MeanAbsoluteErrorSpark
 * (P:69-86, P:256-258; call sites predict/Baseline.scala:46-67, distributed/DistributedBaseline.scala:46,
 * predict/Personalized.scala:61-67, predict/kNN.scala:43-44) ---- */
MRS_API int32_t mrs_predict(const mrs_model* m, const mrs_sim* sim_or_null, int32_t pred_kind,
                            const int32_t* users, const int32_t* items, int64_t n, double* out);
MRS_API int32_t mrs_mae(const mrs_model* m, const mrs_sim* sim_or_null, int32_t pred_kind, const mrs_ratings* test, double* mae_out);
/* asynchronous form: writes {sum |r - p|, count} as two fp64 to device memory (a sharded run adds them across ranks) */
MRS_API int32_t mrs_mae_async(const mrs_model* m, const mrs_sim* sim_or_null, int32_t pred_kind, const mrs_ratings* test,
                              void* device_out2);

/* ---- several GPUs from ONE process (what a single JVM can drive; distributed/DistributedBaseline.scala:30-47 with
 * --master local[N] is one process over N partitions).  mrs_multi_create takes the CUDA device ids, enables peer access
 * between all pairs and creates one engine per device; mrs_multi_load shards both rating sets by user (contiguous id
 * ranges balanced by train rating count) and builds one train set, test set, model and exchange object per device;
 * mrs_multi_baseline_mae runs fit_local on every device, the per-item exchange (our NVLink peer-memory kernel; the peers
 * are plain pointers inside one process), fit_finish, the fused MAE over each device's test pairs and the 16-byte
 * exchange -- every launch asynchronous on its device's stream -- and returns
 * MeanAbsoluteErrorSpark(baselinePredictorSpark(train), test).  mrs_multi_model lends the model of a device slot for the
 * query entry points (item vectors and the global average are global after a pass; a user's average lives on the slot
 * mrs_multi_owner names). ---- */
MRS_API int32_t mrs_multi_create(const int32_t* device_ids, int32_t n_devices, mrs_multi** out);
MRS_API int32_t mrs_multi_load(mrs_multi* m, const int32_t* train_users, const int32_t* train_items, const double* train_ratings, int64_t n_train,
                               const int32_t* test_users, const int32_t* test_items, const double* test_ratings, int64_t n_test);
MRS_API int32_t mrs_multi_baseline_mae(mrs_multi* m, double* mae_out);
MRS_API int32_t mrs_multi_model(mrs_multi* m, int32_t slot, mrs_model** out);
MRS_API int32_t mrs_multi_owner(const mrs_multi* m, int32_t user, int32_t* slot_out);
MRS_API void mrs_multi_destroy(mrs_multi* m);

/* The whole timed closure of distributed/DistributedBaseline.scala:45-47 / predict/Baseline.scala:66-69,
 * MeanAbsoluteErrorSpark(baselinePredictorSpark(train), test), as one asynchronous call on an existing model of `train` with
 * item averages switched off: user sums, item pass and a test pass that finishes the fit in its own prologue (one kernel
 * and one dependency less than mrs_fit_async + mrs_mae_async; same results, same model arrays afterwards). */
MRS_API int32_t mrs_fit_mae_async(mrs_engine* e, const mrs_ratings* train, mrs_model** inout, const mrs_ratings* test, void* device_out2);

/* The same closure for ONE RANK of a user-sharded run (distributed/DistributedBaseline.scala:30-47 over N partitions): both
 * exchanges of the pass happen inside the test pass kernel.  Its CTAs first deliver this rank's per-item partial sums
 * (P:267-268 reduceByKey/collect, P:247 sum/count) into every rank's receive buffer with NVLink stores -- an equal share per
 * CTA -- then wait for all ranks' flags, build their tile's item deviations from the deliveries in their OWN memory in rank
 * order (bit-identical on every rank) and write the model's arrays; the last block exchanges {sum |err|, n} (P:256-258).
 * Three kernels per step and no separate exchange or finishing kernel.  x_items (capacity >= 2*n_known+2 doubles) and
 * x_pair (>= 2) are two connected exchange handles; device_known_items lists, ascending, the items that occur on SOME rank. */
MRS_API int32_t mrs_fit_mae_push_async(mrs_engine* e, const mrs_ratings* train, mrs_model** inout, const mrs_ratings* test,
                                       mrs_exchange* x_items, mrs_exchange* x_pair, const int32_t* device_known_items, int32_t n_known,
                                       void* device_out2);

/* ---- recommendations (P:651-674; call site recommend/Recommender.scala:82-88) ---- */
MRS_API int32_t mrs_recommend(const mrs_model* m, const mrs_sim* sim_or_null, int32_t pred_kind, int32_t user, int32_t n,
                              int32_t* items_out, double* scores_out, int32_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* MRS_B200_H */
