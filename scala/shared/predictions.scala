package shared

import org.apache.spark.rdd.RDD
import org.apache.spark.sql.SparkSession

/**
 * Drop-in replacement of the reference's `package object shared.predictions`
 * (EloDoyard/movie-recommender-system, src/main/scala/shared/predictions.scala, "P:" below): the same public names and
 * signatures, so predict.Baseline, distributed.DistributedBaseline, predict.Personalized, predict.kNN and
 * recommend.Recommender compile unchanged -- but every factory returns a TAGGED function object that carries native
 * handles, and MAE / MeanAbsoluteErrorSpark / recommendations / getNeighbors answer with ONE batched call into
 * libmrs_b200.so (JNI: the Java holder java/shared/ + jni/mrs_jni.c; C ABI: include/mrs_b200.h).
 *
 * Rules kept from the reference: pure functions of immutable rating collections; "missing" ids take the fallbacks of the
 * reference (global average, 0.0), never an error.  Differences: a function that was NOT made by this package cannot
 * run on the GPU, and there is no CPU fallback: passing one raises UnsupportedOperationException.
 *
 * Fitted state is cached per rating-collection IDENTITY (the reference's collections are immutable), so a chain like
 *   predictor(train, weightedSumDeviation(train, getSimilarity(train, 300, adjustedCosineSimilarityFunction(train))))
 * uploads `train` once, fits the baseline model once and computes the similarities once.
 *
 * Scala 2.11 / Java 8 / Spark 2.4.7 like the reference (build.sbt:5-10).  Not compiled in the authoring image (no JDK);
 * the executable mirror with the same structure is movie-recommender-system_b200/predictions.py, and
 * tests/test_jni_shim.py checks that every NativeEngine method used here exists in the Java holder and the JNI shim.
 */
package object predictions {

  case class Rating(user: Int, item: Int, rating: Double)                                        // P:9

  // ------------------------------------------------------------------ timing helpers (P:11-33), unchanged semantics
  def timingInMs(f: () => Double): (Double, Double) = {                                          // P:11
    val start = System.nanoTime()
    val output = f()
    val end = System.nanoTime()
    (output, (end - start) / 1000000.0)
  }
  def mean(s: Seq[Double]): Double = if (s.size > 0) s.reduce(_ + _) / s.length else 0.0          // P:18
  def std(s: Seq[Double]): Double = {                                                            // P:19 (population)
    if (s.size == 0) 0.0
    else { val m = mean(s); scala.math.sqrt(s.map(x => scala.math.pow(m - x, 2)).sum / s.length.toDouble) }
  }
  def toInt(s: String): Option[Int] = try { Some(s.toInt) } catch { case _: Exception => None }  // P:27

  // ------------------------------------------------------------------ device-resident state, cached by identity
  private lazy val engine: Long = NativeEngine.engineCreate(sys.env.getOrElse("MRS_DEVICE", "0").toInt)

  /** One rating collection on the device + what has been fitted on it so far. */
  final class DeviceSet private[predictions] (val ratings: Long) {
    private var modelHandle = 0L
    private val sims = scala.collection.mutable.Map.empty[Int, Long]     // similarity kind -> handle (full sorted rows: any k)
    def model: Long = synchronized { if (modelHandle == 0L) modelHandle = NativeEngine.fit(engine, ratings); modelHandle }
    def sim(kind: Int, k: Int): Long = synchronized {
      if (kind == NativeEngine.SIM_UNIFORM) sims.getOrElseUpdate(kind, NativeEngine.fitSimilarity(model, kind, 0))
      else sims.get(kind) match {
        case Some(h) => h
        case None => val h = NativeEngine.fitSimilarity(model, kind, k); sims(kind) = h; h
      }
    }
    /** the handle of `kind` with neighbourhood size k selected.  Full sorted rows serve every k (the first k of ONE sorted
     *  list, P:610); above 16,384 users only the first k_fit neighbours are kept, and a larger k needs a new fit. */
    def simWithK(kind: Int, k: Int): Long = synchronized {
      val h = sim(kind, k)
      if (kind != NativeEngine.SIM_UNIFORM) {
        try NativeEngine.simSetK(h, k)
        catch { case _: RuntimeException =>
          NativeEngine.simDestroy(h)
          val h2 = NativeEngine.fitSimilarity(model, kind, k); sims(kind) = h2; return h2 }
      }
      h
    }
  }
  private val sets = new java.util.IdentityHashMap[AnyRef, DeviceSet]()

  private def packCodes(rs: Seq[Rating]): Option[Array[Byte]] = {        // half-star data travels as 1 byte per rating
    val out = new Array[Byte](rs.length); var k = 0
    val it = rs.iterator
    while (it.hasNext) { val c = it.next().rating * 2.0; if (c < 0.0 || c > 254.0 || c != c.floor) return None; out(k) = c.toInt.toByte; k += 1 }
    Some(out)
  }
  private def device(rs: Seq[Rating]): DeviceSet = sets.synchronized {
    var d = sets.get(rs)
    if (d == null) {
      val users = new Array[Int](rs.length); val items = new Array[Int](rs.length)
      var k = 0; val it = rs.iterator
      while (it.hasNext) { val r = it.next(); users(k) = r.user; items(k) = r.item; k += 1 }
      val h = packCodes(rs) match {
        case Some(codes) => NativeEngine.ratingsFromCodes(engine, users, items, codes)
        case None => NativeEngine.ratingsFromCoo(engine, users, items, rs.map(_.rating).toArray)
      }
      d = new DeviceSet(h); sets.put(rs, d)
    }
    d
  }
  /** RDDs are materialised once on the driver (no Spark executors on the hot path) and cached by RDD identity. */
  private def device(rdd: RDD[Rating]): DeviceSet = sets.synchronized {
    var d = sets.get(rdd)
    if (d == null) { d = device(rdd.collect().toSeq); sets.put(rdd, d) }
    d
  }

  // ------------------------------------------------------------------ tagged function objects
  final class GpuPredictor private[predictions] (val set: DeviceSet, val kind: Int, val sim: GpuSimilarity) extends ((Int, Int) => Double) {
    private[predictions] def simHandle: Long = if (sim == null) 0L else sim.bind()
    def apply(u: Int, i: Int): Double = NativeEngine.predictOne(set.model, simHandle, kind, u, i)
  }
  /** (u, v) => similarity; k > 0 keeps only the first k neighbours of u (getSimilarity, P:626). */
  final class GpuSimilarity private[predictions] (val set: DeviceSet, val kind: Int, val k: Int) extends ((Int, Int) => Double) {
    /** handle with THIS object's k selected (k is state of the shared native handle: set before every use) */
    private[predictions] def bind(): Long = set.simWithK(kind, k)
    def apply(u: Int, v: Int): Double = if (kind == NativeEngine.SIM_UNIFORM) 1.0 else NativeEngine.similarity(bind(), u, v)
    private[predictions] def on(other: DeviceSet): GpuSimilarity = if (other eq set) this else new GpuSimilarity(other, kind, k)
  }
  final class GpuWsd private[predictions] (val set: DeviceSet, val sim: GpuSimilarity) extends ((Int, Int) => Double) {
    def apply(u: Int, i: Int): Double = NativeEngine.predictOne(set.model, sim.bind(), NativeEngine.PRED_WSD, u, i)
  }

  private def unsupported(what: String) = new UnsupportedOperationException(
    s"$what must be a function made by shared.predictions: arbitrary closures cannot run on the GPU and there is no CPU fallback")
  private def asPredictor(f: (Int, Int) => Double): GpuPredictor = f match { case p: GpuPredictor => p; case _ => throw unsupported("predict") }
  private def asSimilarity(f: (Int, Int) => Double, on: DeviceSet): GpuSimilarity = f match {
    case s: GpuSimilarity => if (s.set == null) new GpuSimilarity(on, s.kind, s.k) else s.on(on)
    case _ => throw unsupported("similarityFunction")
  }

  /** Map[Int, Double] view of a model vector: only ids that occur in the train set are members (P:113, P:134, P:176). */
  private def vectorMap(set: DeviceSet, kind: Int): Map[Int, Double] = {
    val n = NativeEngine.modelVectorSize(set.model, kind)
    val vals = new Array[Double](n); val counts = new Array[Int](n)
    NativeEngine.modelVector(set.model, kind, vals, counts)
    (0 until n).iterator.filter(counts(_) > 0).map(i => i -> vals(i)).toMap
  }
  private def entryMap(set: DeviceSet, which: Int): Map[(Int, Int), Double] = {
    val h = set.sim(NativeEngine.SIM_COSINE, 0)
    val n = NativeEngine.entryValues(h, which, null, null, null).toInt
    val u = new Array[Int](n); val i = new Array[Int](n); val v = new Array[Double](n)
    NativeEngine.entryValues(h, which, u, i, v)
    (0 until n).iterator.map(k => (u(k), i(k)) -> v(k)).toMap
  }

  // ------------------------------------------------------------------ loader (P:35-49)
  def load(spark: SparkSession, path: String, sep: String): RDD[Rating] = {
    val file = spark.sparkContext.textFile(path)
    file.map(l => {
      val cols = l.split(sep).map(_.trim)
      toInt(cols(0)) match {
        case Some(_) => Some(Rating(cols(0).toInt, cols(1).toInt, cols(2).toDouble))
        case None => None
      }
    }).filter(_.isDefined).map(_.get)
  }

  // ------------------------------------------------------------------ baseline family (P:57-237)
  def scale(x: Double, y: Double): Double = if (x > y) 5 - y else if (x < y) y - 1 else 1        // P:57
  def MAE(predict: (Int, Int) => Double, data: Seq[Rating]): Double = {                          // P:69: ONE fused native call
    val p = asPredictor(predict)
    NativeEngine.mae(p.set.model, p.simHandle, p.kind, device(data).ratings)
  }
  def average(ratings: Seq[Rating]): Double = NativeEngine.globalAvg(device(ratings).model)      // P:94
  def computeAvgRating(ratings: Seq[Rating]): (Int, Int) => Double = new GpuPredictor(device(ratings), NativeEngine.PRED_GLOBAL, null)   // P:101
  def usersAvg(ratings: Seq[Rating]): Map[Int, Double] = vectorMap(device(ratings), NativeEngine.VEC_USER)                              // P:113
  def computeUserAvg(ratings: Seq[Rating]): (Int, Int) => Double = new GpuPredictor(device(ratings), NativeEngine.PRED_USER, null)       // P:120
  def itemsAvg(ratings: Seq[Rating]): Map[Int, Double] = vectorMap(device(ratings), NativeEngine.VEC_ITEM)                              // P:134
  def computeItemAvg(ratings: Seq[Rating]): (Int, Int) => Double = new GpuPredictor(device(ratings), NativeEngine.PRED_ITEM, null)       // P:141
  def computeNormalizeDeviation(ratings: Seq[Rating]): Map[(Int, Int), Double] = entryMap(device(ratings), 0)                           // P:155
  def itemsAvgDev(ratings: Seq[Rating]): Map[Int, Double] = vectorMap(device(ratings), NativeEngine.VEC_ITEM_DEV)                       // P:176
  def computeItemAvgDev(ratings: Seq[Rating]): (Int, Int) => Double = new GpuPredictor(device(ratings), NativeEngine.PRED_ITEMDEV, null) // P:193
  def computePrediction(ratings: Seq[Rating]): (Int, Int) => Double = new GpuPredictor(device(ratings), NativeEngine.PRED_BASELINE, null) // P:205

  // ------------------------------------------------------------------ Spark twins (P:246-391): the same engine, fed from the driver
  def meanSpark(r: RDD[Double]): Double = r.sum / r.count                                        // P:246
  def MeanAbsoluteErrorSpark(predictor: (Int, Int) => Double, real: RDD[Rating]): Double = {     // P:256
    val p = asPredictor(predictor)
    NativeEngine.mae(p.set.model, p.simHandle, p.kind, device(real).ratings)
  }
  def getGlobalAvg(r: RDD[Rating]): Double = NativeEngine.globalAvg(device(r).model)                                                    // P:265
  def getUsersAvg(r: RDD[Rating]): Map[Int, Double] = vectorMap(device(r), NativeEngine.VEC_USER)                                       // P:274
  def usersAvgSpark(r: RDD[Rating]): (Int, Int) => Double = new GpuPredictor(device(r), NativeEngine.PRED_USER, null)                   // P:281
  def getItemsAvg(r: RDD[Rating]): Map[Int, Double] = vectorMap(device(r), NativeEngine.VEC_ITEM)                                       // P:295
  def itemsAvgSpark(r: RDD[Rating]): (Int, Int) => Double = new GpuPredictor(device(r), NativeEngine.PRED_ITEM, null)                   // P:302
  def getNormalizedDev(r: RDD[Rating]): Map[(Int, Int), Double] = entryMap(device(r), 0)                                                // P:316
  def getItemsAvgDev(r: RDD[Rating]): Map[Int, Double] = vectorMap(device(r), NativeEngine.VEC_ITEM_DEV)                                // P:336
  def itemsAvgDevSpark(r: RDD[Rating]): (Int, Int) => Double = new GpuPredictor(device(r), NativeEngine.PRED_ITEMDEV, null)             // P:350
  def baselinePredictorSpark(r: RDD[Rating]): (Int, Int) => Double = new GpuPredictor(device(r), NativeEngine.PRED_BASELINE, null)      // P:362

  // ------------------------------------------------------------------ personalized / kNN (P:400-649)
  def similarityOne: (Int, Int) => Double = new GpuSimilarity(null, NativeEngine.SIM_UNIFORM, 0)                                        // P:400
  def adjustedCosineSimilarityFunction(ratings: Seq[Rating]): (Int, Int) => Double = {                                                  // P:407
    val d = device(ratings); d.sim(NativeEngine.SIM_COSINE, 0); new GpuSimilarity(d, NativeEngine.SIM_COSINE, 0)
  }
  def jaccardCoefficient(ratings: Seq[Rating]): (Int, Int) => Double = {                                                                // P:440
    val d = device(ratings); d.sim(NativeEngine.SIM_JACCARD, 0); new GpuSimilarity(d, NativeEngine.SIM_JACCARD, 0)
  }
  def preprocessedRating(ratings: Seq[Rating]): Map[(Int, Int), Double] = entryMap(device(ratings), 1)                                  // P:470
  def weightedSumDeviation(ratings: Seq[Rating], similarityFunction: (Int, Int) => Double): (Int, Int) => Double = {                    // P:489
    val d = device(ratings); new GpuWsd(d, asSimilarity(similarityFunction, d))
  }
  def predictor(ratings: Seq[Rating], weightedSumDevFunction: (Int, Int) => Double): (Int, Int) => Double = {                           // P:557
    val d = device(ratings)
    weightedSumDevFunction match {
      case w: GpuWsd => new GpuPredictor(d, NativeEngine.PRED_PERSONALIZED, w.sim.on(d))
      case _ => throw unsupported("weightedSumDevFunction")
    }
  }
  def getNeighbors(ratings: Seq[Rating], k: Int, similarityFunction: (Int, Int) => Double): Int => Seq[(Int, Double)] = {               // P:596
    val d = device(ratings); val s = new GpuSimilarity(d, asSimilarity(similarityFunction, d).kind, k)
    (u: Int) => {
      val ids = new Array[Int](scala.math.max(k, 1)); val sims = new Array[Double](scala.math.max(k, 1))
      val n = NativeEngine.neighbors(s.bind(), u, k, ids, sims)
      (0 until n).map(j => (ids(j), sims(j)))
    }
  }
  def getSimilarity(ratings: Seq[Rating], k: Int, similarityFunction: (Int, Int) => Double): (Int, Int) => Double = {                   // P:626
    val d = device(ratings); new GpuSimilarity(d, asSimilarity(similarityFunction, d).kind, k)
  }

  // ------------------------------------------------------------------ recommendations (P:651-674)
  def recommendations(ratings: Seq[Rating], predictor: (Int, Int) => Double): (Int, Int) => Seq[(Int, Double)] = {
    val p = asPredictor(predictor)
    (user: Int, n: Int) => {
      val items = new Array[Int](scala.math.max(n, 1)); val scores = new Array[Double](scala.math.max(n, 1))
      val w = NativeEngine.recommend(p.set.model, p.simHandle, p.kind, user, n, items, scores)
      (0 until w).map(j => (items(j), scores(j)))
    }
  }

  // ------------------------------------------------------------------ several GPUs from this JVM (distributed.DistributedBaseline with --master local[N])
  /** MeanAbsoluteErrorSpark(baselinePredictorSpark(train), test) with the users sharded over `devices` (one process, NVLink exchange). */
  def distributedBaselineMae(devices: Seq[Int], train: RDD[Rating], test: RDD[Rating]): Double = {
    val tr = train.collect(); val te = test.collect()
    val multi = NativeEngine.multiCreate(devices.toArray)
    try {
      NativeEngine.multiLoad(multi, tr.map(_.user), tr.map(_.item), tr.map(_.rating), te.map(_.user), te.map(_.item), te.map(_.rating))
      NativeEngine.multiBaselineMae(multi)
    } finally NativeEngine.multiDestroy(multi)
  }
}
