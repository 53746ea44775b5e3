#!/bin/bash
# exploratory: the reference's eltwise / equation drivers (prebuilt, unmodified) on the GPU; prints rc and the verdict line of each
cd tests/c/_drivers || exit 1
export LD_LIBRARY_PATH=$PWD/../../../libxsmm_b200/lib:$LD_LIBRARY_PATH OMP_NUM_THREADS=4
run() { name=$1; shift; out=$(timeout -s KILL 60 ./$name "$@" 2>&1); rc=$?; echo "[$rc] $name $* :: $(echo "$out" | grep -i -E 'success|fail|error|correct' | tail -2 | tr '\n' '|' | cut -c1-200)"; }
for op in 1 2 3 5 11; do run eltwise_unary_simple $op 0 F32 F32 F32 37 11 40 40 0; done
run eltwise_unary_simple 1 0 BF16 F32 BF16 64 16 64 64 0
run eltwise_unary_simple 1 0 F32 F32 BF8 64 16 64 64 0
run eltwise_unary_simple 1 0 F32 F32 BF8 64 16 64 64 1
run eltwise_unary_simple 1 1 F32 F32 F32 32 32 32 32 0
for op in 1 2 3 4; do run eltwise_binary_simple $op 0 F32 F32 F32 F32 37 11 40 40; done
run eltwise_binary_simple 1 3 BF16 BF16 F32 BF16 64 16 64 64
run eltwise_ternary_simple 1 0 F32 F32 F32 F32 F32 37 11 40 40
run eltwise_ternary_simple 3 0 BF16 BF16 BF16 F32 BF16 64 16 64 64
run eltwise_unary_dropout F 1 F32 F32 64 16 64 64
run eltwise_unary_dropout B 1 F32 F32 64 16 64 64
run eltwise_unary_dropout F 0 BF16 BF16 33 7 40 40
run eltwise_unary_gather_scatter 64 32 80 64 0 0 0 0 1
run eltwise_unary_gather_scatter 64 32 64 80 1 1 1 1 1
run eltwise_unary_quantization F32 I8 64 16 64 64 0 0
run eltwise_unary_quantization F32 I16 33 7 40 40 0 1
run eltwise_unary_quantization_to_mxbf8 64 16 64 64
run eltwise_unary_quantization_to_mxfp4 64 16 64 64
run eltwise_unary_quantization_to_nvfp4 64 16 64 64
run eltwise_unary_reduce 64 32 64 1 0 0 0 F32 0 0 0 0 1
run eltwise_unary_reduce 64 32 64 1 1 1 0 F32 0 0 0 0 1
run eltwise_unary_reduce 64 32 64 1 0 0 1 F32 0 0 1 0 1
run equation_simple 64 32
run equation_relu 64 32
run equation_softmax 64 32
run gimmik
run gemm_kernel_parallel F32 F32 F32 F32 64 64 64 64 64 64 1 1 0 0 0 0 0 0 nobr 1 1 100
