#!/bin/bash
# The C++ readers of the library (SPIR-V disassembler, SPIR-V -> IR, IR text, IR -> PTX) under AddressSanitizer and
# UndefinedBehaviorSanitizer on mutated inputs; no GPU involved.   bash tools/asan_fuzz/run.sh [seeds...]
# Last run (end of round 2, with the version 2 translator and readers): 3 seeds x 24000 inputs, see the commit message.
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
OUT=${TMPDIR:-/tmp}/kzn_asan_fuzz
mkdir -p $OUT
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -I$ROOT/kazan_b200/csrc \
    $ROOT/tools/asan_fuzz/driver.cpp $ROOT/kazan_b200/csrc/ir_text.cpp $ROOT/kazan_b200/csrc/spirv_dis.cpp \
    $ROOT/kazan_b200/csrc/spirv_to_ir.cpp $ROOT/kazan_b200/csrc/spirv_to_ir_v2.cpp $ROOT/kazan_b200/csrc/ptx_emit.cpp -o $OUT/driver
for seed in ${@:-1 2 3}; do  # (timeout: an endless loop in a reader is a finding too)
  python $ROOT/tools/asan_fuzz/make_corpus.py $seed 4000 $OUT/corpus_$seed.bin
  timeout 900 $OUT/driver $OUT/corpus_$seed.bin
done
