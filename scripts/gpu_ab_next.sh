#!/bin/bash
# A/B candidates prepared at the end of round 1 (not yet measured on a GPU).  Build the variant libraries first:
#   bash scripts/build_ab_next.sh
# then:  gpurun --timeout 1500 -- bash scripts/gpu_ab_next.sh <tag>      (~15 min of box time; every step has its own timeout)
set -u
OUT=gpurun_out/${1:-ab_next}; mkdir -p $OUT
B=$PWD/build_ab
rk() { MARLPDE_B200_LIB=$1 MARLPDE_PROFILE_EVENTS=1 timeout 120 python scripts/profile_rk45.py 300 5 > $OUT/rk45_$2.log 2>&1; echo "rk45 $2: $(tail -2 $OUT/rk45_$2.log | tr '\n' ' ')"; }
IN=$PWD/integrating-diagenetic-equations-using-python_b200/marlpde_b200/libmarlpde_b200.so
rk $IN intree
rk $B/lib_imm.so imm          # -DMARLPDE_FP64_IMM=1: constants with a zero low word as instruction immediates (results change in the last bits)
rk $IN intree_again
# MARLPDE_RK45_BUILD=450: experimental 4-cells-per-thread / 8-warp kernel (csrc/rk45_quad.cu); every run under
# a short timeout: the kernel has never been on a GPU
MARLPDE_RK45_BUILD=450 timeout 60 python scripts/profile_rk45.py 300 5 > $OUT/rk45_quad.log 2>&1; echo "rk45 quad (no events): $(tail -2 $OUT/rk45_quad.log | tr '\n' ' ')"
MARLPDE_RK45_BUILD=450 MARLPDE_PROFILE_EVENTS=1 timeout 60 python scripts/profile_rk45.py 300 5 > $OUT/rk45_quad_ev.log 2>&1; echo "rk45 quad: $(tail -2 $OUT/rk45_quad_ev.log | tr '\n' ' ')"
MARLPDE_B200_LIB=$B/lib_quad_o1.so MARLPDE_RK45_BUILD=450 MARLPDE_PROFILE_EVENTS=1 timeout 60 python scripts/profile_rk45.py 300 5 > $OUT/rk45_quad_o1.log 2>&1; echo "rk45 quad, own parts of both pairs before the wait: $(tail -2 $OUT/rk45_quad_o1.log | tr '\n' ' ')"
MARLPDE_B200_LIB=$B/lib_quad_rolled.so MARLPDE_RK45_BUILD=450 MARLPDE_PROFILE_EVENTS=1 timeout 60 python scripts/profile_rk45.py 300 5 > $OUT/rk45_quad_rolled.log 2>&1; echo "rk45 quad, one RHS instance in a rolled pair loop (half the hot code, spills): $(tail -2 $OUT/rk45_quad_rolled.log | tr '\n' ' ')"
for m in 2 3; do   # other RHS instruction schedules under the 8-warp kernel (255 registers change what ptxas can overlap)
  MARLPDE_B200_LIB=$B/lib_m$m.so MARLPDE_RK45_BUILD=450 MARLPDE_PROFILE_EVENTS=1 timeout 60 python scripts/profile_rk45.py 300 5 > $OUT/rk45_quad_m$m.log 2>&1; echo "rk45 quad, RHS schedule $m: $(tail -1 $OUT/rk45_quad_m$m.log)"
done
MARLPDE_RK45_BUILD=450 timeout 300 python -m pytest -q -x -m gpu -p no:cacheprovider --timeout=60 --timeout-method=thread tests/test_gpu_rk45.py tests/test_gpu_dropin.py > $OUT/pytest_quad.log 2>&1
echo "pytest quad: $(tail -3 $OUT/pytest_quad.log | tr '\n' ' ')"
# r01i first contact: build 450 is correct but only +3.6 % (static estimate +25 %): capture it, after the plain runs above exited
MARLPDE_RK45_BUILD=450 MARLPDE_PROFILE_EVENTS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rk45_quad -s 1 -c 1 -o $OUT/rk45_quad_full python scripts/profile_rk45.py 300 3 > $OUT/ncu_quad_full.log 2>&1; echo "ncu quad: rc $?"
# -DMARLPDE_TAIL_SPREAD=1: slot s of a CTA claims only while more than s * gridDim columns are left, grid = min(SMs, columns): the
# last (partly filled) round of a sweep and batches smaller than the machine run one column per SM (results are slot independent:
# bit-identical).  4096 columns x 3000 attempts (the bench step: 9.2 rounds of 444 slots), 64 columns, build 450 too; then parity.
for lat in 16,16,16 4,4,4; do
  for l in $IN $B/lib_spread.so; do
    MARLPDE_PROFILE_LATTICE=$lat MARLPDE_B200_LIB=$l MARLPDE_PROFILE_EVENTS=1 timeout 120 python scripts/profile_rk45.py 3000 3 > $OUT/rk45_spread_tmp.log 2>&1
    echo "rk45 lattice $lat $(basename $l): $(tail -1 $OUT/rk45_spread_tmp.log)"; cat $OUT/rk45_spread_tmp.log >> $OUT/rk45_spread.log
  done
done
MARLPDE_RK45_BUILD=450 MARLPDE_B200_LIB=$B/lib_spread.so MARLPDE_PROFILE_EVENTS=1 timeout 120 python scripts/profile_rk45.py 3000 3 > $OUT/rk45_quad_spread.log 2>&1; echo "rk45 quad + spread: $(tail -1 $OUT/rk45_quad_spread.log)"
MARLPDE_B200_LIB=$B/lib_spread.so timeout 400 python -m pytest -q -x -m gpu -p no:cacheprovider --timeout=120 --timeout-method=thread tests/test_gpu_rk45.py tests/test_gpu_dropin.py > $OUT/pytest_spread.log 2>&1
echo "pytest spread: $(tail -1 $OUT/pytest_spread.log)"
# -DMARLPDE_TILE_TMA=1: the tile kernel of the large-N path moves its windows with 1-D TMA bulk copies (UBLKCP; never run on a GPU):
# parity first (streaming tests, under a short timeout), then N = 20 000 x 64 and N = 2 000 x 64 against the in-tree library
MARLPDE_B200_LIB=$B/lib_tma.so timeout 300 python -m pytest -q -x -m gpu -p no:cacheprovider --timeout=60 --timeout-method=thread tests/test_gpu_rk45.py -k streaming > $OUT/pytest_tma.log 2>&1
echo "pytest tma: $(tail -1 $OUT/pytest_tma.log)"
for l in $IN $B/lib_tma.so; do
  for n in 20000 2000; do
    MARLPDE_B200_LIB=$l timeout 120 python scripts/profile_stream.py $n 64 > $OUT/stream_tmp.log 2>&1; echo "stream N=$n x 64 $(basename $l): $(tail -1 $OUT/stream_tmp.log)"
  done
done
# MARLPDE_RK45_STREAM_GRAPH=1 (runtime switch of the in-tree library): a batch of the streaming path captured once into a CUDA graph and
# replayed; small batches are launch bound (N = 20 000 x 1: 37 k attempts/s).  Parity (streaming tests), then timing with 256 attempts.
MARLPDE_RK45_STREAM_GRAPH=1 timeout 300 python -m pytest -q -x -m gpu -p no:cacheprovider --timeout=60 --timeout-method=thread tests/test_gpu_rk45.py -k streaming > $OUT/pytest_graph.log 2>&1
echo "pytest graph: $(tail -1 $OUT/pytest_graph.log)"
for g in 0 1; do
  for nb in "20000 1" "2000 8" "20000 64"; do
    MARLPDE_RK45_STREAM_GRAPH=$g timeout 120 python scripts/profile_stream.py $nb 256 > $OUT/stream_tmp.log 2>&1; echo "stream $nb graph=$g: $(tail -1 $OUT/stream_tmp.log)"
  done
done
MARLPDE_B200_LIB=$B/lib_imm.so timeout 120 python scripts/profile_radau.py 16 0.05 > $OUT/radau_imm.log 2>&1; echo "radau imm: $(head -1 $OUT/radau_imm.log)"
timeout 120 python scripts/profile_radau.py 16 0.05 > $OUT/radau_intree.log 2>&1; echo "radau in-tree: $(head -1 $OUT/radau_intree.log)"
# -DMARLPDE_RADAU_FUSE_F=1: the three stage evaluations of a Newton iteration fused with B = TI F - M W (6 x 5N fewer doubles through DRAM)
MARLPDE_B200_LIB=$B/lib_fuse.so timeout 120 python scripts/profile_radau.py 16 0.05 > $OUT/radau_fuse.log 2>&1; echo "radau fuse: $(head -1 $OUT/radau_fuse.log)"
MARLPDE_B200_LIB=$B/lib_fuse.so timeout 120 python scripts/profile_radau.py 4 0.05 > $OUT/radau_fuse_64.log 2>&1; echo "radau fuse, 64 columns: $(head -1 $OUT/radau_fuse_64.log)"
timeout 120 python scripts/profile_radau.py 4 0.05 > $OUT/radau_intree_64.log 2>&1; echo "radau in-tree, 64 columns: $(head -1 $OUT/radau_intree_64.log)"
MARLPDE_B200_LIB=$B/lib_fuse.so timeout 400 python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread tests/test_gpu_radau.py tests/test_gpu_dropin.py > $OUT/pytest_fuse.log 2>&1
echo "pytest fuse: $(tail -1 $OUT/pytest_fuse.log)"
MARLPDE_B200_LIB=$B/lib_imm.so timeout 600 python -m pytest -q -m gpu -p no:cacheprovider --timeout=300 --timeout-method=thread tests > $OUT/pytest_imm.log 2>&1
echo "pytest imm: $(tail -1 $OUT/pytest_imm.log)"
echo done
