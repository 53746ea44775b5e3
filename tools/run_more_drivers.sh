#!/bin/bash
# exploratory: the reference's eltwise / equation drivers (prebuilt, unmodified) on the GPU; prints rc and the verdict line of each
cd tests/c/_drivers || exit 1
export LD_LIBRARY_PATH=$PWD/../../../libxsmm_b200/lib:$LD_LIBRARY_PATH OMP_NUM_THREADS=4
run() { name=$1; shift; out=$(timeout -s KILL 60 ./$name "$@" 2>&1); rc=$?; echo "[$rc] $name $* :: $(echo "$out" | grep -i -E 'success|fail|error|correct|Total Max Error' | tail -2 | tr '\n' '|' | cut -c1-200)"; }
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
run equation_simple_layernorm
run equation_bf16_x3_split_f32 32 16 40
run gimmik 3
# the drivers that so far ran only against the simulated device (tests/test_hostsim.py holds the full lists)
gk() { run "$@"; }
gk gemm_kernel F32 F32 F32 F32 64 64 64 64 64 64 1 0 0 0 0 0 0 0 0 nopf nobr 1 0 3 0
gk gemm_kernel F32 F32 F32 F32 37 21 45 40 48 40 1 1 0 0 0 1 0 0 0 nopf strdbr 4 0 3 0
gk gemm_kernel F32 F32 F32 F32 32 32 32 32 32 32 1 1 0 0 0 0 0 0 0 nopf addrbr 3 0 3 0
gk gemm_kernel F32 F32 F32 F32 32 32 32 32 32 32 1 0 0 0 0 0 0 0 0 nopf offsbr 3 0 3 0
gk gemm_kernel BF16 BF16 F32 BF16 64 64 64 64 64 64 1 0 0 0 0 0 1 0 0 nopf strdbr 2 0 3 0
gk gemm_kernel BF16 BF16 F32 BF16 64 64 64 64 64 64 1 0 0 0 0 0 1 0 1 nopf nobr 1 0 3 0
gk gemm_kernel I8 I8 I32 I32 64 64 64 64 64 64 1 0 0 0 0 0 1 0 0 nopf nobr 1 0 3 0
gk gemm_kernel BF8 BF8 F32 F32 64 64 64 64 64 64 1 1 0 0 0 0 1 0 0 nopf nobr 1 0 3 0
gk gemm_kernel F32 F32 F32 F32 64 64 64 64 64 64 1 0 0 0 0 0 0 0 0 nopf spmm 4 0 3 0
for b in 0 1; do for u in 0 1 2 3; do gk gemm_kernel_fused F32 F32 F32 F32 64 48 32 64 32 64 1 1 0 0 0 0 0 0 0 nopf nobr 1 0 3 0 $b $u; done; done
gk gemm_kernel_fused BF16 BF16 F32 BF16 64 64 64 64 64 64 1 0 0 0 0 0 1 0 1 nopf nobr 1 0 3 0 1 1
gk gemm_kernel_parallel F32 F32 F32 F32 64 64 64 64 64 64 1 1 0 0 0 0 0 0 0 nopf nobr 1 0 3 0
for t in D L E; do run eltwise_unary_relu $t F 1 F32 F32 F32 37 11 48 40; run eltwise_unary_relu $t B 1 F32 F32 F32 37 11 48 40; done
for op in T R S V W Q F G H I X Y Z B C D; do run eltwise_unary_transform $op BF16 32 16 32 32; done
run eltwise_ternary_simple 1 0 F32 F32 IMPLICIT F32 F32 64 16 64 64
run eltwise_ternary_simple 1 0 F32 F32 IMPLICIT F32 BF8 64 16 64 64 1
for u in ut_threadsafety ut_registry ut_gemmflags ut_matdiff; do run $u; done
