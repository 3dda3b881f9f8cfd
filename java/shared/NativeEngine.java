package shared;

/**
 * Java 8 holder of the native methods implemented in jni/mrs_jni.c over libmrs_b200.so (include/mrs_b200.h).
 * Handles are opaque longs; every failure of the native library surfaces as a RuntimeException with its message.
 * The Scala facade scala/shared/predictions.scala is the only intended caller.
 *
 * Kinds (include/mrs_b200.h):
 *   vecKind  0 global average, 1 user averages, 2 item averages, 3 item average deviations
 *   predKind 0 global, 1 user, 2 item, 3 item deviation, 4 baseline, 5 personalized (needs a similarity), 6 weighted-sum deviation
 *   simKind  0 uniform (similarityOne), 1 adjusted cosine, 2 Jaccard
 */
public final class NativeEngine {
  static { System.loadLibrary("mrs_jni"); }
  private NativeEngine() {}

  public static final int VEC_GLOBAL = 0, VEC_USER = 1, VEC_ITEM = 2, VEC_ITEM_DEV = 3;
  public static final int PRED_GLOBAL = 0, PRED_USER = 1, PRED_ITEM = 2, PRED_ITEMDEV = 3, PRED_BASELINE = 4, PRED_PERSONALIZED = 5, PRED_WSD = 6;
  public static final int SIM_UNIFORM = 0, SIM_COSINE = 1, SIM_JACCARD = 2;

  public static native long engineCreate(int device);
  public static native void engineDestroy(long engine);
  public static native String version();

  public static native long ratingsFromCoo(long engine, int[] users, int[] items, double[] ratings);
  /** codes[k] = (byte) (2 * rating): the compact form for half-star data (9 bytes per rating over PCIe instead of 16). */
  public static native long ratingsFromCodes(long engine, int[] users, int[] items, byte[] codes);
  public static native long ratingsFromText(long engine, byte[] text, String sep);
  public static native long ratingsFromFile(long engine, String path, String sep);
  public static native long ratingsCount(long ratings);
  public static native void ratingsDestroy(long ratings);

  public static native long fit(long engine, long train);
  public static native void modelDestroy(long model);
  public static native double globalAvg(long model);
  public static native double modelLookup(long model, int vecKind, int id);
  public static native int modelVectorSize(long model, int vecKind);
  public static native void modelVector(long model, int vecKind, double[] valsOut, int[] countsOut);

  public static native long fitSimilarity(long model, int simKind, int k);
  public static native void simSetK(long sim, int k);
  public static native double similarity(long sim, int u, int v);
  public static native int neighbors(long sim, int u, int k, int[] idsOut, double[] simsOut);
  /** usersOut == null: returns the number of entries; else fills the three arrays (user-major order). */
  public static native long entryValues(long sim, int which, int[] usersOut, int[] itemsOut, double[] valsOut);
  public static native void simDestroy(long sim);

  public static native double predictOne(long model, long simOrZero, int predKind, int user, int item);
  public static native void predict(long model, long simOrZero, int predKind, int[] users, int[] items, double[] out);
  public static native double mae(long model, long simOrZero, int predKind, long test);
  public static native int recommend(long model, long simOrZero, int predKind, int user, int n, int[] itemsOut, double[] scoresOut);

  public static native long multiCreate(int[] deviceIds);
  public static native void multiLoad(long multi, int[] trainUsers, int[] trainItems, double[] trainRatings,
                                      int[] testUsers, int[] testItems, double[] testRatings);
  public static native double multiBaselineMae(long multi);
  public static native long multiModel(long multi, int slot);
  public static native int multiOwner(long multi, int user);
  public static native void multiDestroy(long multi);
}
