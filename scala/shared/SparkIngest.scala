package shared

import org.apache.spark.rdd.RDD
import shared.predictions.Rating

/**
 * Spark-side ingestion shim (SURVEY 8(f).4): feeds an RDD[Rating] to the engine WITHOUT collecting 20 M Rating objects
 * into one Seq on the driver (what `load(...).collect()` at predict/Baseline.scala:40-42 or the RDD closures of
 * distributed/DistributedBaseline.scala:41-43 would do).  Every partition is turned into three primitive arrays on its
 * executor (mapPartitions), the driver pulls the partitions one at a time (toLocalIterator: at most one partition of
 * primitives in flight) and appends them to growable primitive buffers -- in the compact form (int, int, byte code =
 * 2 x rating) when the data are half-star, i.e. 9 bytes per rating of driver memory and of PCIe traffic instead of a
 * 32-byte object plus a pointer.  The result is a device-resident rating set handle for NativeEngine.
 *
 * Source only: needs Spark 2.4.7 / Scala 2.11 to compile (not present in the authoring image).
 */
object SparkIngest {
  final case class Chunk(users: Array[Int], items: Array[Int], ratings: Array[Double])

  def chunks(rdd: RDD[Rating]): RDD[Chunk] = rdd.mapPartitions { it =>
    val u = new scala.collection.mutable.ArrayBuilder.ofInt
    val i = new scala.collection.mutable.ArrayBuilder.ofInt
    val r = new scala.collection.mutable.ArrayBuilder.ofDouble
    it.foreach { x => u += x.user; i += x.item; r += x.rating }
    Iterator.single(Chunk(u.result(), i.result(), r.result()))
  }

  /** Device-resident rating set of `rdd` on `engine` (a NativeEngine handle); the caller owns the returned handle. */
  def toDevice(engine: Long, rdd: RDD[Rating]): Long = {
    val u = new scala.collection.mutable.ArrayBuilder.ofInt
    val i = new scala.collection.mutable.ArrayBuilder.ofInt
    val r = new scala.collection.mutable.ArrayBuilder.ofDouble
    val c = new scala.collection.mutable.ArrayBuilder.ofByte
    var halfStar = true
    chunks(rdd).toLocalIterator.foreach { ch =>
      u ++= ch.users; i ++= ch.items; r ++= ch.ratings
      if (halfStar) ch.ratings.foreach { x =>
        val k = x * 2.0
        if (halfStar && k >= 0.0 && k <= 254.0 && k == k.floor) c += k.toInt.toByte else halfStar = false
      }
    }
    if (halfStar) NativeEngine.ratingsFromCodes(engine, u.result(), i.result(), c.result())
    else NativeEngine.ratingsFromCoo(engine, u.result(), i.result(), r.result())
  }
}
