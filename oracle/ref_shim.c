/* TEST INFRASTRUCTURE -- not part of the product.
 *
 * Builds the UNMODIFIED reference (LIBXSMM, /root/reference) in its header-only mode into
 * oracle/_ref/libxsmm_ref.so and exposes a small plain-C interface ("ref_*") over
 *   (1) the reference's portable C kernels  libxsmm_reference_gemm / libxsmm_reference_elementwise
 *       (src/generator_gemm_reference_impl.c:2818, src/generator_mateltwise_reference_impl.c:2663),
 *   (2) the reference's own JIT path (AMX/AVX-512 on this host) through its public dispatch API,
 *   (3) OpenMP batch drivers over (2) used as the CPU baseline of bench.py --impl reference.
 * No reference source is copied: this file only #includes it from where it lies. The recipe is
 * `make ref` (one gcc call). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the resulting library.
 */
#include <libxsmm_source.h>
#include <omp.h>
#include <string.h>

#define REF_API __attribute__((visibility("default")))

REF_API const char* ref_target_arch(void) { libxsmm_init(); return libxsmm_get_target_arch(); }
REF_API int ref_max_threads(void) { return omp_get_max_threads(); }

static libxsmm_gemm_batch_reduce_config ref_brcfg(int br_type, long long sa, long long sb) {
  return libxsmm_create_gemm_batch_reduce_config(
    br_type == 1 ? LIBXSMM_GEMM_BATCH_REDUCE_ADDRESS : (br_type == 2 ? LIBXSMM_GEMM_BATCH_REDUCE_OFFSET
    : (br_type == 3 ? LIBXSMM_GEMM_BATCH_REDUCE_STRIDE : LIBXSMM_GEMM_BATCH_REDUCE_NONE)), (libxsmm_blasint)sa, (libxsmm_blasint)sb, 0);
}

static void ref_fill_param(libxsmm_gemm_param* p, unsigned long long* br, void* a, void* b, void* c,
                           long long* offs_a, long long* offs_b, float* scf) {
  memset(p, 0, sizeof(*p));
  p->op.tertiary = br; p->a.primary = a; p->b.primary = b; p->c.primary = c;
  p->a.secondary = offs_a; p->b.secondary = offs_b; p->c.tertiary = scf;
}

/* dims = {m,n,k,lda,ldb,ldc}; types = {a,b,comp,c}; mode 0: C reference kernel, 1: JIT kernel.
 * returns 0 ok, 1 dispatch failed, 2 JIT fell back to the reference kernel (still executed) */
REF_API int ref_gemm(const int* dims, const int* types, unsigned int flags, int br_type, long long stride_a, long long stride_b,
                     unsigned long long br, void* a, void* b, void* c, long long* offs_a, long long* offs_b, float scf, int mode)
{
  const libxsmm_gemm_shape shape = libxsmm_create_gemm_shape(dims[0], dims[1], dims[2], dims[3], dims[4], dims[5],
    (libxsmm_datatype)types[0], (libxsmm_datatype)types[1], (libxsmm_datatype)types[3], (libxsmm_datatype)types[2]);
  const libxsmm_gemm_batch_reduce_config cfg = ref_brcfg(br_type, stride_a, stride_b);
  libxsmm_gemm_param p;
  unsigned long long brv = br;
  libxsmm_init();
  ref_fill_param(&p, &brv, a, b, c, offs_a, offs_b, &scf);
  if (mode == 0) {
    libxsmm_descriptor_blob blob;
    const libxsmm_gemm_descriptor* desc = libxsmm_gemm_descriptor_init_brgemm(&blob, shape, flags, 0, cfg);
    if (desc == NULL) return 1;
    libxsmm_reference_gemm(&p, desc);
    return 0;
  } else {
    libxsmm_xmmfunction k; libxsmm_kernel_info info;
    k.gemm = libxsmm_dispatch_brgemm(shape, flags, 0, cfg);
    if (k.gemm == NULL) return 1;
    k.gemm(&p);
    libxsmm_get_kernel_info(k.ptr_const, &info);
    return info.is_reference_kernel ? 2 : 0;
  }
}

/* matrix equation through the reference (JIT): nodes in pre-order, 8 ints each {type 1 arg / 2 unary / 3 binary / 4 ternary, op, dtype,
 * flags, pos, m, n, ld}; out = {m, n, ld, type}; inputs[] = the argument matrices */
REF_API int ref_meqn(const int* nodes, int nnodes, const int* out, void** inputs, int ninputs, void* output)
{
  libxsmm_blasint eq; int i; libxsmm_meqn_function f; libxsmm_meqn_param p; libxsmm_matrix_arg args[16];
  libxsmm_init();
  eq = libxsmm_meqn_create();
  for (i = 0; i < nnodes; ++i) {
    const int* nd = nodes + 8 * i;
    if (nd[0] == 1) libxsmm_meqn_push_back_arg(libxsmm_create_meqn_arg_metadata(eq, nd[4]), libxsmm_create_meqn_arg_shape(nd[5], nd[6], nd[7], (libxsmm_datatype)nd[2]),
                                               libxsmm_create_matrix_arg_attributes(LIBXSMM_MATRIX_ARG_TYPE_SINGULAR, LIBXSMM_MATRIX_ARG_SET_TYPE_NONE, 0, 0));
    else if (nd[0] == 2) libxsmm_meqn_push_back_unary_op(libxsmm_create_meqn_op_metadata(eq, nd[4]), (libxsmm_meltw_unary_type)nd[1], (libxsmm_datatype)nd[2], (libxsmm_bitfield)nd[3]);
    else if (nd[0] == 3) libxsmm_meqn_push_back_binary_op(libxsmm_create_meqn_op_metadata(eq, nd[4]), (libxsmm_meltw_binary_type)nd[1], (libxsmm_datatype)nd[2], (libxsmm_bitfield)nd[3]);
    else libxsmm_meqn_push_back_ternary_op(libxsmm_create_meqn_op_metadata(eq, nd[4]), (libxsmm_meltw_ternary_type)nd[1], (libxsmm_datatype)nd[2], (libxsmm_bitfield)nd[3]);
  }
  f = libxsmm_dispatch_meqn(eq, libxsmm_create_meqn_arg_shape(out[0], out[1], out[2], (libxsmm_datatype)out[3]));
  if (f == NULL || ninputs > 16) return 1;
  memset(&p, 0, sizeof(p)); memset(args, 0, sizeof(args));
  for (i = 0; i < ninputs; ++i) args[i].primary = inputs[i];
  p.inputs = args; p.output.primary = output;
  f(&p);
  return 0;
}

/* packed dense GEMM through the reference's JIT; kind 0: libxsmm_create_packed_gemm, 1: _ac_rm, 2: _bc_rm */
REF_API int ref_packed_dense(int kind, int dtype, const int* dims, unsigned int flags, int packed_width, void* a, void* b, void* c)
{
  const libxsmm_gemm_shape shape = libxsmm_create_gemm_shape(dims[0], dims[1], dims[2], dims[3], dims[4], dims[5],
    (libxsmm_datatype)dtype, (libxsmm_datatype)dtype, (libxsmm_datatype)dtype, (libxsmm_datatype)dtype);
  libxsmm_gemm_param p; libxsmm_gemmfunction k;
  libxsmm_init();
  k = (kind == 0) ? libxsmm_create_packed_gemm(shape, flags, 0, packed_width)
    : ((kind == 1) ? libxsmm_create_packed_gemm_ac_rm(shape, flags, 0, packed_width) : libxsmm_create_packed_gemm_bc_rm(shape, flags, 0, packed_width));
  if (k == NULL) return 1;
  memset(&p, 0, sizeof(p));
  p.a.primary = a; p.b.primary = b; p.c.primary = c;
  k(&p);
  libxsmm_release_kernel((const void*)k);
  return 0;
}

/* aux_kind 1: int4 A with zero points (aux -> a.quaternary, flags completed like samples/xgemm/gemm_kernel.c:2892-2900);
 * aux_kind 2: bitmap-compressed A (aux -> a.secondary). C reference kernel only. */
REF_API int ref_gemm_aux(const int* dims, const int* types, unsigned int flags, int br_type, long long stride_a, long long stride_b,
                         unsigned long long br, void* a, void* b, void* c, int aux_kind, void* aux)
{
  const libxsmm_gemm_shape shape = libxsmm_create_gemm_shape(dims[0], dims[1], dims[2], dims[3], dims[4], dims[5],
    (libxsmm_datatype)types[0], (libxsmm_datatype)types[1], (libxsmm_datatype)types[3], (libxsmm_datatype)types[2]);
  const libxsmm_gemm_batch_reduce_config cfg = ref_brcfg(br_type, stride_a, stride_b);
  libxsmm_gemm_param p; unsigned long long brv = br; libxsmm_descriptor_blob blob; const libxsmm_gemm_descriptor* desc;
  libxsmm_init();
  memset(&p, 0, sizeof(p));
  p.op.tertiary = &brv; p.a.primary = a; p.b.primary = b; p.c.primary = c;
  if (aux_kind == 1) p.a.quaternary = aux; else p.a.secondary = aux;
  desc = libxsmm_gemm_descriptor_init_brgemm(&blob, shape, flags, 0, cfg);
  if (desc == NULL) return 1;
  libxsmm_reference_gemm(&p, desc);
  return 0;
}

/* fused form: fuse = {colbias, cp_op (0 / RELU / SIGMOID), relu bitmask, vnni_c}; same calling convention as oracle_gemm_ext */
REF_API int ref_gemm_ext(const int* dims, const int* types, unsigned int flags, int br_type, long long stride_a, long long stride_b,
                         unsigned long long br, void* a, void* b, void* c, long long* offs_a, long long* offs_b, float scf,
                         const int* fuse, void* colbias, unsigned char* relu_mask, int mode)
{
  const libxsmm_gemm_shape shape = libxsmm_create_gemm_shape(dims[0], dims[1], dims[2], dims[3], dims[4], dims[5],
    (libxsmm_datatype)types[0], (libxsmm_datatype)types[1], (libxsmm_datatype)types[3], (libxsmm_datatype)types[2]);
  const libxsmm_gemm_batch_reduce_config cfg = ref_brcfg(br_type, stride_a, stride_b);
  const libxsmm_gemm_ext_unary_argops argops = libxsmm_create_gemm_ext_unary_argops(0, LIBXSMM_MELTW_TYPE_UNARY_NONE, LIBXSMM_MELTW_FLAG_UNARY_NONE, 0,
    0, LIBXSMM_MELTW_TYPE_UNARY_NONE, LIBXSMM_MELTW_FLAG_UNARY_NONE, 0,
    dims[5], (libxsmm_meltw_unary_type)fuse[1], fuse[2] ? LIBXSMM_MELTW_FLAG_UNARY_BITMASK_2BYTEMULT : LIBXSMM_MELTW_FLAG_UNARY_NONE, 0);
  const libxsmm_gemm_ext_binary_postops postops = libxsmm_create_gemm_ext_binary_postops(dims[5], (libxsmm_datatype)types[3],
    fuse[0] ? LIBXSMM_MELTW_TYPE_BINARY_ADD : LIBXSMM_MELTW_TYPE_BINARY_NONE, fuse[0] ? LIBXSMM_MELTW_FLAG_BINARY_BCAST_COL_IN_0 : LIBXSMM_MELTW_FLAG_BINARY_NONE);
  libxsmm_gemm_ext_param p; unsigned long long brv = br;
  const unsigned int fl = flags | (fuse[3] ? LIBXSMM_GEMM_FLAG_VNNI_C : 0);
  libxsmm_init();
  memset(&p, 0, sizeof(p));
  p.op.tertiary = &brv; p.a.primary = a; p.b.primary = b; p.c.primary = c; p.a.secondary = offs_a; p.b.secondary = offs_b; p.c.tertiary = &scf;
  p.d.primary = colbias; p.c.secondary = relu_mask;
  if (mode == 0) {
    libxsmm_descriptor_blob blob;
    const libxsmm_gemm_descriptor* desc = libxsmm_gemm_descriptor_init_brgemm_ext(&blob, shape, fl, 0, cfg, argops, postops);
    if (desc == NULL) return 1;
    libxsmm_reference_gemm(&p, desc);
    return 0;
  } else {
    libxsmm_gemmfunction_ext k = libxsmm_dispatch_brgemm_ext(shape, fl, 0, cfg, argops, postops);
    if (k == NULL) return 1;
    k(&p);
    return 0;
  }
}

/* desc = {op_class, op, flags, m, n, ldi, ldi2, ldi3, ldo, t_in0, t_in1, t_in2, t_out, t_comp};
 * param points to a libxsmm_meltw_{unary,binary,ternary}_param image. mode as above. */
REF_API int ref_meltw(const int* desc, void* param, int mode) {
  libxsmm_descriptor_blob blob;
  const libxsmm_meltw_descriptor* d;
  libxsmm_init();
  d = libxsmm_meltw_descriptor_init2(&blob, (libxsmm_datatype)desc[9], (libxsmm_datatype)desc[10], (libxsmm_datatype)desc[11],
        (libxsmm_datatype)desc[13], (libxsmm_datatype)desc[12], desc[3], desc[4], desc[5], desc[8], desc[6], desc[7],
        (unsigned short)desc[2], (unsigned short)desc[1], (unsigned char)desc[0]);
  if (mode == 0) { libxsmm_reference_elementwise(param, d); return 0; }
  else {
    libxsmm_xmeltwfunction k = libxsmm_dispatch_meltw(d);
    if (k.xmeltw == NULL) return 1;
    k.xmeltw(param);
    return 0;
  }
}

REF_API int ref_fsspmdm(int dtype, int M, int N, int K, int lda, int ldb, int ldc, const void* alpha, const void* beta,
                        const void* a_dense, const void* B, void* C)
{
  libxsmm_fsspmdm* h = libxsmm_fsspmdm_create((libxsmm_datatype)dtype, M, N, K, lda, ldb, ldc, alpha, beta, a_dense, 0, NULL);
  if (h == NULL) return 1;
  libxsmm_fsspmdm_execute(h, B, C);
  libxsmm_fsspmdm_destroy(h);
  return 0;
}

/* types = {a,b,comp,c}; geometry = {m_blocks, M(packed width), K, N, bk, bn} */
REF_API int ref_bcsc(const int* types, const int* geo, unsigned int flags, void* A, void* Bvals, unsigned int* colptr,
                     unsigned int* rowidx, void* C)
{
  const libxsmm_gemm_shape shape = libxsmm_create_gemm_shape(geo[0], 0, geo[2], geo[2], 0, geo[3],
    (libxsmm_datatype)types[0], (libxsmm_datatype)types[1], (libxsmm_datatype)types[3], (libxsmm_datatype)types[2]);
  libxsmm_spgemm_config cfg; libxsmm_gemm_param p; libxsmm_gemmfunction k;
  unsigned long long nblk = (unsigned long long)(geo[3] / geo[5]);
  libxsmm_init();
  cfg.packed_width = geo[1]; cfg.bk = geo[4]; cfg.bn = geo[5];
  k = libxsmm_create_packed_spgemm_bcsc(shape, flags, 0, cfg);
  if (k == NULL) return 1;
  memset(&p, 0, sizeof(p));
  p.a.primary = A; p.b.primary = Bvals; p.b.secondary = colptr; p.b.tertiary = rowidx; p.b.quaternary = &nblk; p.c.primary = C;
  k(&p);
  libxsmm_release_kernel((const void*)k);
  return 0;
}

/* is_csc: 0 -> create_packed_spgemm_csr, 1 -> _csc; dims = {m,n,k,lda,ldb,ldc} */
REF_API int ref_packed_sp(int is_csc, int dtype, const int* dims, unsigned int flags, int packed_width,
                          const unsigned int* ptr, const unsigned int* idx, const void* values, void* a, void* b, void* c)
{
  const libxsmm_gemm_shape shape = libxsmm_create_gemm_shape(dims[0], dims[1], dims[2], dims[3], dims[4], dims[5],
    (libxsmm_datatype)dtype, (libxsmm_datatype)dtype, (libxsmm_datatype)dtype, (libxsmm_datatype)dtype);
  libxsmm_gemm_param p; libxsmm_gemmfunction k;
  libxsmm_init();
  k = is_csc ? libxsmm_create_packed_spgemm_csc(shape, flags, 0, packed_width, ptr, idx, values)
             : libxsmm_create_packed_spgemm_csr(shape, flags, 0, packed_width, ptr, idx, values);
  if (k == NULL) return 1;
  memset(&p, 0, sizeof(p));
  p.a.primary = a; p.b.primary = b; p.c.primary = c;
  k(&p);
  libxsmm_release_kernel((const void*)k);
  return 0;
}

/* scalar conversions, for pinning the oracle's and the kernels' rounding rules */
REF_API unsigned short ref_f32_to_bf16(float f) { return libxsmm_convert_f32_to_bf16_rne(f); }
REF_API unsigned short ref_f32_to_f16(float f) { return libxsmm_convert_f32_to_f16(f); }
REF_API float ref_f16_to_f32(unsigned short h) { return libxsmm_convert_f16_to_f32(h); }
REF_API float ref_bf16_to_f32(unsigned short h) { return libxsmm_convert_bf16_to_f32(h); }

/* ---- CPU baseline drivers: the reference's JIT kernel, batch loop over all host cores -------------- */
/* strided batch of count tiles (same convention as libxsmm_b200_gemm_batch_strided); returns seconds
 * for `reps` passes, negative on dispatch failure. is_ref[0] tells whether the C fallback was used. */
REF_API double ref_bench_gemm_batch(const int* dims, const int* types, unsigned int flags, int br_type, long long stride_a,
  long long stride_b, unsigned long long br, char* a, char* b, char* c, long long ta, long long tb, long long tc,
  long long count, int reps, int* is_ref)
{
  const libxsmm_gemm_shape shape = libxsmm_create_gemm_shape(dims[0], dims[1], dims[2], dims[3], dims[4], dims[5],
    (libxsmm_datatype)types[0], (libxsmm_datatype)types[1], (libxsmm_datatype)types[3], (libxsmm_datatype)types[2]);
  const libxsmm_gemm_batch_reduce_config cfg = ref_brcfg(br_type, stride_a, stride_b);
  libxsmm_xmmfunction k; libxsmm_kernel_info info;
  libxsmm_timer_tickint t0; int rep;
  libxsmm_init();
  k.gemm = libxsmm_dispatch_brgemm(shape, flags, 0, cfg);
  if (k.gemm == NULL) return -1.0;
  libxsmm_get_kernel_info(k.ptr_const, &info);
  if (is_ref != NULL) *is_ref = (int)info.is_reference_kernel;
  t0 = libxsmm_timer_tick();
  for (rep = 0; rep < reps; ++rep) {
    long long t;
#   pragma omp parallel for schedule(static)
    for (t = 0; t < count; ++t) {
      libxsmm_gemm_param p; unsigned long long brv = br;
      memset(&p, 0, sizeof(p));
      p.op.tertiary = &brv; p.a.primary = a + t * ta; p.b.primary = b + t * tb; p.c.primary = c + t * tc;
      k.gemm(&p);
    }
  }
  return libxsmm_timer_duration(t0, libxsmm_timer_tick());
}

/* The reference arm of bench.py: the SAME strided batch as the GPU arm (count tiles, all operands unique), owned by this
 * function so that every page is first touched by the thread that will stream it (NUMA-local like
 * samples/xgemm/gemm_kernel_parallel.c), filled once, then `passes` timed passes (seconds per pass in out_seconds).
 * Values: multiples of 0.1 in [-0.5, 0.5] as bf16 (the drivers' fill, spmm_kernel.c:498-527). Returns 0, or -1 when
 * dispatch fails, -2 when the host cannot hold the buffers. */
REF_API int ref_bench_brgemm_owned(int m, int n, int k, unsigned long long br, unsigned int flags, long long count, int warm,
                                   int passes, double* out_seconds, int* is_ref, double* checksum)
{
  static const unsigned short tenth_bf16[11] = { 0xBF00, 0xBECD, 0xBE9A, 0xBE4D, 0xBDCD, 0x0000, 0x3DCD, 0x3E4D, 0x3E9A, 0x3ECD, 0x3F00 };
  const libxsmm_gemm_shape shape = libxsmm_create_gemm_shape(m, n, k, m, k, m, LIBXSMM_DATATYPE_BF16, LIBXSMM_DATATYPE_BF16,
    LIBXSMM_DATATYPE_F32, LIBXSMM_DATATYPE_F32);
  const long long ta = (long long)br * m * k, tb = (long long)br * k * n, tc = (long long)m * n;   /* elements per tile */
  const libxsmm_gemm_batch_reduce_config cfg = ref_brcfg(3, (long long)m * k * 2, (long long)k * n * 2);
  libxsmm_xmmfunction kern; libxsmm_kernel_info info;
  unsigned short *a, *b; float* c; long long t; int pass; double sum = 0;
  libxsmm_init();
  kern.gemm = libxsmm_dispatch_brgemm(shape, flags, 0, cfg);
  if (kern.gemm == NULL) return -1;
  libxsmm_get_kernel_info(kern.ptr_const, &info);
  if (is_ref != NULL) *is_ref = (int)info.is_reference_kernel;
  a = (unsigned short*)malloc((size_t)count * ta * 2); b = (unsigned short*)malloc((size_t)count * tb * 2); c = (float*)malloc((size_t)count * tc * 4);
  if (a == NULL || b == NULL || c == NULL) { free(a); free(b); free(c); return -2; }
# pragma omp parallel for schedule(static)
  for (t = 0; t < count; ++t) {
    unsigned int x = 555u + (unsigned int)t * 2654435761u; long long i;
    for (i = 0; i < ta; ++i) { x = x * 1664525u + 1013904223u; a[t * ta + i] = tenth_bf16[(x >> 16) % 11u]; }
    for (i = 0; i < tb; ++i) { x = x * 1664525u + 1013904223u; b[t * tb + i] = tenth_bf16[(x >> 16) % 11u]; }
    for (i = 0; i < tc; ++i) c[t * tc + i] = 0.f;
  }
  for (pass = -warm; pass < passes; ++pass) {
    const libxsmm_timer_tickint t0 = libxsmm_timer_tick();
#   pragma omp parallel for schedule(static)
    for (t = 0; t < count; ++t) {
      libxsmm_gemm_param p; unsigned long long brv = br;
      memset(&p, 0, sizeof(p));
      p.op.tertiary = &brv; p.a.primary = a + t * ta; p.b.primary = b + t * tb; p.c.primary = c + t * tc;
      kern.gemm(&p);
    }
    if (pass >= 0) out_seconds[pass] = libxsmm_timer_duration(t0, libxsmm_timer_tick());
  }
  for (t = 0; t < count; t += 997) sum += c[t * tc + (t % tc)];
  if (checksum != NULL) *checksum = sum;
  free(a); free(b); free(c);
  return 0;
}

/* fsspmdm: N split into one slice per thread like samples/xgemm_sparse_Ainregs/pyfr_driver_asp_reg.c:380-394 */
REF_API double ref_bench_fsspmdm(int dtype, int M, int N, int K, int lda, const void* alpha, const void* beta,
                                 const void* a_dense, const char* B, char* C, int reps)
{
  const int nthreads = omp_get_max_threads();
  const int ts = (dtype == LIBXSMM_DATATYPE_F64) ? 8 : 4, vl = 64 / ts;
  int nslice = (N / vl) / nthreads * vl, rep;
  libxsmm_fsspmdm *h_main, *h_tail = NULL;
  libxsmm_timer_tickint t0;
  libxsmm_init();
  if (nslice <= 0) nslice = N;
  h_main = libxsmm_fsspmdm_create((libxsmm_datatype)dtype, M, nslice, K, lda, N, N, alpha, beta, a_dense, 0, NULL);
  if (h_main == NULL) return -1.0;
  if (N % nslice != 0) h_tail = libxsmm_fsspmdm_create((libxsmm_datatype)dtype, M, N % nslice, K, lda, N, N, alpha, beta, a_dense, 0, NULL);
  t0 = libxsmm_timer_tick();
  for (rep = 0; rep < reps; ++rep) {
    const int nfull = N / nslice; int s;
#   pragma omp parallel for schedule(static)
    for (s = 0; s < nfull; ++s) libxsmm_fsspmdm_execute(h_main, B + (size_t)s * nslice * ts, C + (size_t)s * nslice * ts);
    if (h_tail != NULL) libxsmm_fsspmdm_execute(h_tail, B + (size_t)nfull * nslice * ts, C + (size_t)nfull * nslice * ts);
  }
  {
    const double dt = libxsmm_timer_duration(t0, libxsmm_timer_tick());
    libxsmm_fsspmdm_destroy(h_main); if (h_tail != NULL) libxsmm_fsspmdm_destroy(h_tail);
    return dt;
  }
}

/* BCSC: m_blocks split into contiguous ranges, one per thread */
REF_API double ref_bench_bcsc(const int* types, const int* geo, unsigned int flags, char* A, void* Bvals, unsigned int* colptr,
                              unsigned int* rowidx, char* C, int reps)
{
  const int nthreads = omp_get_max_threads();
  const int per = (geo[0] + nthreads - 1) / nthreads;
  const size_t tsa = LIBXSMM_TYPESIZE((libxsmm_datatype)types[0]), tsc = LIBXSMM_TYPESIZE((libxsmm_datatype)types[3]);
  libxsmm_spgemm_config cfg; libxsmm_gemmfunction k_main, k_tail = NULL;
  unsigned long long nblk = (unsigned long long)(geo[3] / geo[5]);
  libxsmm_timer_tickint t0; int rep;
  const int nfull = geo[0] / per, tail = geo[0] % per;
  libxsmm_init();
  cfg.packed_width = geo[1]; cfg.bk = geo[4]; cfg.bn = geo[5];
  k_main = libxsmm_create_packed_spgemm_bcsc(libxsmm_create_gemm_shape(per, 0, geo[2], geo[2], 0, geo[3],
    (libxsmm_datatype)types[0], (libxsmm_datatype)types[1], (libxsmm_datatype)types[3], (libxsmm_datatype)types[2]), flags, 0, cfg);
  if (k_main == NULL) return -1.0;
  if (tail > 0) k_tail = libxsmm_create_packed_spgemm_bcsc(libxsmm_create_gemm_shape(tail, 0, geo[2], geo[2], 0, geo[3],
    (libxsmm_datatype)types[0], (libxsmm_datatype)types[1], (libxsmm_datatype)types[3], (libxsmm_datatype)types[2]), flags, 0, cfg);
  t0 = libxsmm_timer_tick();
  for (rep = 0; rep < reps; ++rep) {
    int s;
#   pragma omp parallel for schedule(static)
    for (s = 0; s < nfull + (tail > 0 ? 1 : 0); ++s) {
      libxsmm_gemm_param p; unsigned long long nb = nblk;
      memset(&p, 0, sizeof(p));
      p.a.primary = A + (size_t)s * per * geo[2] * geo[1] * tsa; p.b.primary = Bvals; p.b.secondary = colptr; p.b.tertiary = rowidx;
      p.b.quaternary = &nb; p.c.primary = C + (size_t)s * per * geo[3] * geo[1] * tsc;
      if (s < nfull) k_main(&p); else k_tail(&p);
    }
  }
  return libxsmm_timer_duration(t0, libxsmm_timer_tick());
}

/* ---- utilities (libxsmm_utils.h): reference values for tests/test_host_utils.py ------------------------ */
/* out[0..21] = the 22 doubles of libxsmm_matdiff_info in declaration order, out[22..25] = m, n, i, r, out[26] = epsilon */
REF_API int ref_matdiff(int dtype, int m, int n, const void* ref, const void* tst, int ldr, int ldt, double* out) {
  libxsmm_matdiff_info d;
  const libxsmm_blasint lr = ldr, lt = ldt;
  const int rc = libxsmm_matdiff(&d, (libxsmm_datatype)dtype, m, n, ref, tst, ldr > 0 ? &lr : NULL, ldt > 0 ? &lt : NULL);
  memcpy(out, &d, 22 * sizeof(double));
  out[22] = d.m; out[23] = d.n; out[24] = d.i; out[25] = d.r; out[26] = libxsmm_matdiff_epsilon(&d);
  return rc;
}
/* folds `count` comparisons of column slices with libxsmm_matdiff_reduce; same output image */
REF_API int ref_matdiff_reduce(int dtype, int m, int n, int count, const void* ref, const void* tst, double* out) {
  libxsmm_matdiff_info total, d; int i;
  const size_t ts = (size_t)LIBXSMM_TYPESIZE((libxsmm_datatype)dtype);
  libxsmm_matdiff_clear(&total);
  for (i = 0; i < count; ++i) {
    if (0 != libxsmm_matdiff(&d, (libxsmm_datatype)dtype, m, n, (const char*)ref + ts * m * n * i, (const char*)tst + ts * m * n * i, NULL, NULL)) return 1;
    libxsmm_matdiff_reduce(&total, &d);
  }
  memcpy(out, &total, 22 * sizeof(double));
  out[22] = total.m; out[23] = total.n; out[24] = total.i; out[25] = total.r; out[26] = libxsmm_matdiff_epsilon(&total);
  return 0;
}
REF_API unsigned long long ref_coprime2(unsigned long long n) { return (unsigned long long)libxsmm_coprime2((size_t)n); }
REF_API void ref_rng(unsigned int seed, float* f32_seq, int n32, double* f64_seq, int n64, unsigned int* u32_seq, int nu, unsigned int u_range) {
  int i;
  libxsmm_rng_set_seed(seed);
  libxsmm_rng_f32_seq(f32_seq, n32);
  for (i = 0; i < n64; ++i) f64_seq[i] = libxsmm_rng_f64();
  for (i = 0; i < nu; ++i) u32_seq[i] = libxsmm_rng_u32(u_range);
}
/* external generator state, stochastic bf8 conversion and the small math helpers the eltwise drivers call */
REF_API void ref_extstate(unsigned int seed, unsigned int* out64) {
  unsigned int* st = libxsmm_rng_create_extstate(seed);
  memcpy(out64, st, libxsmm_rng_get_extstate_size());
  libxsmm_rng_destroy_extstate(st);
}
REF_API void ref_stochastic_bf8(const float* in, unsigned char* out, unsigned int len, unsigned int* state64, unsigned int start_idx) {
  libxsmm_stochastic_convert_fp32_bf8(in, (libxsmm_bfloat8*)out, len, state64, start_idx);
}
REF_API float ref_sexp2_i8i(int x) { return libxsmm_sexp2_i8i(x); }
REF_API float ref_nearbyintf(float x) { return libxsmm_nearbyintf(x); }
REF_API void ref_lp_convert(int which, const void* in, void* out, unsigned long long n) {
  switch (which) {
    case 0: libxsmm_rne_convert_fp32_bf8((const float*)in, (libxsmm_bfloat8*)out, (size_t)n); break;
    case 1: libxsmm_convert_bf8_f32((const libxsmm_bfloat8*)in, (float*)out, (size_t)n); break;
    case 2: libxsmm_rne_convert_fp32_hf8((const float*)in, (libxsmm_hfloat8*)out, (size_t)n); break;
    case 3: libxsmm_convert_hf8_f32((const libxsmm_hfloat8*)in, (float*)out, (size_t)n); break;
    case 4: libxsmm_rne_convert_fp32_bf16((const float*)in, (libxsmm_bfloat16*)out, (size_t)n); break;
    case 5: libxsmm_rnaz_convert_fp32_bf16((const float*)in, (libxsmm_bfloat16*)out, (size_t)n); break;
    case 6: libxsmm_truncate_convert_f32_bf16((const float*)in, (libxsmm_bfloat16*)out, (size_t)n); break;
    case 7: libxsmm_convert_bf16_f32((const libxsmm_bfloat16*)in, (float*)out, (size_t)n); break;
    case 8: libxsmm_rne_convert_fp32_f16((const float*)in, (libxsmm_float16*)out, (size_t)n); break;
    default: libxsmm_convert_f16_f32((const libxsmm_float16*)in, (float*)out, (size_t)n);
  }
}
/* the fill the drivers use (LIBXSMM_MATINIT), f32 and f64 */
REF_API void ref_matinit(int is_f64, double seed, void* dst, int nrows, int ncols, int ld, double scale) {
  if (is_f64) { LIBXSMM_MATINIT(double, seed, dst, nrows, ncols, ld, scale); }
  else { LIBXSMM_MATINIT(float, seed, dst, nrows, ncols, ld, scale); }
}
