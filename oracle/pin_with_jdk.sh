#!/bin/bash
# pin_with_jdk.sh -- turn "parity unpinned" into a one-command check on any host that has a JDK (>= 8).
#
# The reference's operator file src/sw/SmithWaterman.java imports only three foreign types
# (org.apache.spark.api.java.function.Function2 / Function3 and scala.Tuple2, SmithWaterman.java:3-6), so it
# compiles against three stub classes without Spark or Scala.  This script
#   1. writes the stubs + a small driver into a temp dir,
#   2. compiles the UNMODIFIED reference file from where it lies (never copied into this repo),
#   3. runs the Java operator over every known-answer vector of tests/golden/kat.json and over --random N seeded
#      random pairs (the same generator as oracle/pin_pairs.py prints),
#   4. diffs score / max-cell count / beginnings / both alignment strings against the C oracle's committed answers.
# Exit code 0 = the oracle (and with it every GPU parity test) is pinned to the real Java implementation.
#
#   oracle/pin_with_jdk.sh [/path/to/reference] [--random 200]
#
# NOT RUNNABLE in the build image (no JDK, no network); committed so that a maintainer can.
set -euo pipefail
REF="${1:-/root/reference}"
RANDOM_N=200
if [ "${2:-}" = "--random" ]; then RANDOM_N="${3:-200}"; fi
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
command -v javac >/dev/null || { echo "no javac on PATH: install a JDK (>= 8) first" >&2; exit 2; }
[ -f "$REF/src/sw/SmithWaterman.java" ] || { echo "reference not found at $REF" >&2; exit 2; }
W="$(mktemp -d)"; trap 'rm -rf "$W"' EXIT
mkdir -p "$W/src/org/apache/spark/api/java/function" "$W/src/scala" "$W/src/pin" "$W/out"
cat > "$W/src/org/apache/spark/api/java/function/Function2.java" <<'J'
package org.apache.spark.api.java.function;
public interface Function2<T1, T2, R> extends java.io.Serializable { R call(T1 a, T2 b) throws Exception; }
J
cat > "$W/src/org/apache/spark/api/java/function/Function3.java" <<'J'
package org.apache.spark.api.java.function;
public interface Function3<T1, T2, T3, R> extends java.io.Serializable { R call(T1 a, T2 b, T3 c) throws Exception; }
J
cat > "$W/src/scala/Tuple2.java" <<'J'
package scala;
public class Tuple2<A, B> implements java.io.Serializable {
    private final A a; private final B b;
    public Tuple2(A a, B b) { this.a = a; this.b = b; }
    public A _1() { return a; }
    public B _2() { return b; }
}
J
cat > "$W/src/pin/Pin.java" <<'J'
package pin;
import java.io.*; import java.util.*;
import scala.Tuple2;
/** stdin: one case per line "ref<TAB>read<TAB>match<TAB>mismatch<TAB>gap"; stdout: one line per case
 *  "score<TAB>nCells<TAB>beginning:refAln:readAln;..." (first 64 alignments) from the reference's own operator. */
public class Pin {
    public static void main(String[] a) throws Exception {
        BufferedReader in = new BufferedReader(new InputStreamReader(System.in, "ISO-8859-1"));
        PrintStream out = new PrintStream(new FileOutputStream(FileDescriptor.out), true, "ISO-8859-1");
        char[] types = { 'a', 'i', 'd', '-' };       // Distribution.java:37
        for (String line; (line = in.readLine()) != null; ) {
            String[] f = line.split("\t", -1);
            int[] scores = { Integer.parseInt(f[2]), Integer.parseInt(f[3]), Integer.parseInt(f[4]) };
            Tuple2<Integer, ArrayList<Tuple2<Integer, String[]>>> r =
                new sw.SmithWaterman.OptAlignments().call(new String[]{ f[0], f[1] }, scores, types);
            StringBuilder sb = new StringBuilder();
            sb.append(r._1()).append('\t').append(r._2().size()).append('\t');
            int k = 0;
            for (Tuple2<Integer, String[]> s : r._2()) {
                if (k++ == 64) break;
                sb.append(s._1()).append(':').append(s._2()[0]).append(':').append(s._2()[1]).append(';');
            }
            out.println(sb);
        }
    }
}
J
javac -nowarn -d "$W/out" $(find "$W/src" -name '*.java') "$REF/src/sw/SmithWaterman.java"
python3 "$ROOT/oracle/pin_pairs.py" --random "$RANDOM_N" --emit-cases > "$W/cases.tsv"
java -Xss512m -cp "$W/out" pin.Pin < "$W/cases.tsv" > "$W/java.tsv"
python3 "$ROOT/oracle/pin_pairs.py" --random "$RANDOM_N" --compare "$W/java.tsv"
