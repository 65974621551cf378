#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the *unmodified* reference gap_closer sources where they lie
# (/root/reference/gap_closer) into oracle/_ref/.  Nothing here is shipped or measured as the
# product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
# legs may execute what this script produces.
#
# Outputs (all git-ignored, NOT gpurun-ignored, so they travel to the GPU box):
#   oracle/_ref/gc                   the reference CLI (gap_closer/main.c:120-212)
#   oracle/_ref/ref_kmer             harness: reference chop/put/search/stat with dumps + timings
#   oracle/_ref/libref_sw_asis.so    reference sw.c + cigar.c + hash_func.c verbatim behind a tiny wrapper
#   oracle/_ref/libref_sw_fixed.so   same, with the one-line traceback re-fetch patch
#
# Recipe notes: the shipped `-O3` for every file dead-locks under gcc 13 because
# ont.c spins on a non-volatile flag (ont.c:102-118,359-394).  ont.c is therefore compiled at
# -O1, everything else at -O3.  No reference source is copied into the repository: objects are
# built in a mktemp dir and only binaries land in oracle/_ref/.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF_GAP_CLOSER:-/root/reference/gap_closer}"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "[build_ref] $REF not present (GPU box?) - keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
CC="${CC:-gcc}"

# ---- 1. reference objects (position independent so the same objects serve .so and binaries)
for f in "$REF"/*.c; do
  b="$(basename "$f" .c)"
  O=-O3; [ "$b" = ont ] && O=-O1
  $CC $O -w -fPIC -I"$REF" -c "$f" -o "$TMP/$b.o"
done

# ---- 2. the reference CLI
ALL_OBJS=""
for b in main sw utils hash hash_func str gap_closer kmer contig ont rseq bio gc_graph digraph cigar ctg_graph lfr; do
  ALL_OBJS="$ALL_OBJS $TMP/$b.o"
done
$CC -O3 -o "$OUT/gc" $ALL_OBJS -lm -lz -lpthread

# ---- 3. k-mer harness (everything but main.o)
LIB_OBJS="${ALL_OBJS/$TMP\/main.o/}"
$CC -O2 -w -I"$REF" -c "$HERE/ref_kmer_harness.c" -o "$TMP/ref_kmer_harness.o"
$CC -O2 -o "$OUT/ref_kmer" "$TMP/ref_kmer_harness.o" $LIB_OBJS -lm -lz -lpthread

# ---- 4. SW shared libraries: as-is and fixed (patched copy lives only in $TMP)
$CC -O2 -w -fPIC -I"$REF" -c "$HERE/ref_sw_wrap.c" -o "$TMP/ref_sw_wrap.o"
$CC -shared -o "$OUT/libref_sw_asis.so" "$TMP/ref_sw_wrap.o" "$TMP/sw.o" "$TMP/cigar.o" "$TMP/hash_func.o" "$TMP/utils.o" "$TMP/str.o" -lm -lpthread
# fixed: re-fetch the current cell at the top of the traceback do-body (sw.c:292-293)
sed 's/^\tdo {$/\tdo {\n\t\tc = sm + (bt_tidx<<sw->qry_nbits) + bt_qidx;/' "$REF/sw.c" > "$TMP/sw_fixed.c"
if ! diff <(grep -c 'c = sm + (bt_tidx<<sw->qry_nbits) + bt_qidx;' "$TMP/sw_fixed.c") <(echo 2) >/dev/null; then
  echo "[build_ref] traceback patch did not apply exactly once" >&2; exit 1
fi
$CC -O3 -w -fPIC -I"$REF" -c "$TMP/sw_fixed.c" -o "$TMP/sw_fixed.o"
$CC -shared -o "$OUT/libref_sw_fixed.so" "$TMP/ref_sw_wrap.o" "$TMP/sw_fixed.o" "$TMP/cigar.o" "$TMP/hash_func.o" "$TMP/utils.o" "$TMP/str.o" -lm -lpthread
echo "[build_ref] ok -> $OUT"
