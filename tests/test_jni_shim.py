"""The JNI / Java / Scala binding shipped as source (jni/, java/, scala/): there is no JDK in this image, so the checks are
  * jni/mrs_jni.c compiles (gcc -fsyntax-only -Wall -Werror) against include/mrs_b200.h and a minimal jni.h stand-in, i.e.
    every call into the C ABI has the right arity and types;
  * every `native` method of java/shared/NativeEngine.java has its Java_shared_NativeEngine_* function in the shim and
    vice versa;
  * every NativeEngine member the Scala sources use exists in the Java holder;
  * the Scala facade declares every public function of the reference's package object."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read(*p):
    with open(os.path.join(ROOT, *p)) as f:
        return f.read()


def test_shim_compiles_against_the_c_abi():
    cmd = ["gcc", "-std=c11", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-I", os.path.join(ROOT, "tests", "stubs"),
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "jni", "mrs_jni.c")]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr


def test_java_natives_and_shim_functions_match():
    java = read("java", "shared", "NativeEngine.java")
    natives = set(re.findall(r"public static native [\w\[\]]+ (\w+)\(", java))
    shim = set(re.findall(r"Java_shared_NativeEngine_(\w+)\(", read("jni", "mrs_jni.c")))
    assert natives and natives == shim, (sorted(natives - shim), sorted(shim - natives))


def test_scala_sources_only_use_declared_natives():
    java = read("java", "shared", "NativeEngine.java")
    declared = set(re.findall(r"public static native [\w\[\]]+ (\w+)\(", java)) | set(re.findall(r"\b([A-Z][A-Z_]+) =", java))
    for src in ("predictions.scala", "SparkIngest.scala"):
        used = set(re.findall(r"NativeEngine\.(\w+)", read("scala", "shared", src)))
        assert used <= declared, (src, sorted(used - declared))


def test_facade_declares_the_reference_surface():
    scala = read("scala", "shared", "predictions.scala")
    surface = ["timingInMs", "mean", "std", "toInt", "load", "scale", "MAE", "average", "computeAvgRating", "usersAvg", "computeUserAvg", "itemsAvg",
               "computeItemAvg", "computeNormalizeDeviation", "itemsAvgDev", "computeItemAvgDev", "computePrediction", "meanSpark",
               "MeanAbsoluteErrorSpark", "getGlobalAvg", "getUsersAvg", "usersAvgSpark", "getItemsAvg", "itemsAvgSpark", "getNormalizedDev",
               "getItemsAvgDev", "itemsAvgDevSpark", "baselinePredictorSpark", "similarityOne", "adjustedCosineSimilarityFunction",
               "jaccardCoefficient", "preprocessedRating", "weightedSumDeviation", "predictor", "getNeighbors", "getSimilarity", "recommendations"]
    missing = [f for f in surface if not re.search(r"\bdef %s\b" % f, scala)]
    assert not missing, missing
    assert "case class Rating(user: Int, item: Int, rating: Double)" in scala
