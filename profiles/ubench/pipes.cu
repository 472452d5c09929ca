// Micro-benchmark: issue cost (cycles per warp instruction per SM sub-partition) of the integer instructions the
// arithmetic coder is made of, on sm_100a.  1024 threads per SM (8 warps per scheduler), 4 independent chains per
// thread, so the number measured is throughput, not latency.   nvcc -arch=sm_100a -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITER 4096

#define BODY(NAME, ...)                                                                                    \
  __global__ void __launch_bounds__(1024, 1) NAME(uint32_t *out, uint32_t seed, unsigned long long *cyc)   \
  {                                                                                                        \
    const uint32_t tx = threadIdx.x * 2654435761u; uint32_t a = seed + tx, b = seed * 3 + 1 + (tx >> 3), c = seed * 5 + 2 + (tx >> 5), d = seed * 7 + 3 + (tx >> 7), k = (seed + (tx >> 9)) | 1;   \
    uint32_t e = a ^ 0x55, f = b ^ 0x66, g = c ^ 0x77, h = d ^ 0x88;                                       \
    const unsigned long long t0 = clock64();                                                               \
    _Pragma("unroll 1") for (int i = 0; i < ITER; i++)                                                     \
    {                                                                                                      \
      _Pragma("unroll") for (int j = 0; j < 4; j++) { __VA_ARGS__ }                                               \
    }                                                                                                      \
    const unsigned long long t1 = clock64();                                                               \
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + e + f + g + h;                            \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                                       \
  }

// 8 instructions per STMT
BODY(k_imad, asm volatile("mad.lo.u32 %0, %0, %8, %4; mad.lo.u32 %1, %1, %8, %5; mad.lo.u32 %2, %2, %8, %6; mad.lo.u32 %3, %3, %8, %7;"
                          "mad.lo.u32 %4, %4, %8, %0; mad.lo.u32 %5, %5, %8, %1; mad.lo.u32 %6, %6, %8, %2; mad.lo.u32 %7, %7, %8, %3;"
                          : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k));)
BODY(k_imadhi, asm volatile("mul.hi.u32 %0, %0, %8; mul.hi.u32 %1, %1, %8; mul.hi.u32 %2, %2, %8; mul.hi.u32 %3, %3, %8;"
                            "mul.hi.u32 %4, %4, %8; mul.hi.u32 %5, %5, %8; mul.hi.u32 %6, %6, %8; mul.hi.u32 %7, %7, %8;"
                            : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k));)
BODY(k_imadwide, {
  unsigned long long x0 = ((unsigned long long)a << 32) | b, x1 = ((unsigned long long)c << 32) | d;
  unsigned long long x2 = ((unsigned long long)e << 32) | f, x3 = ((unsigned long long)g << 32) | h;
  asm volatile("mad.wide.u32 %0, %4, %8, %0; mad.wide.u32 %1, %5, %8, %1; mad.wide.u32 %2, %6, %8, %2; mad.wide.u32 %3, %7, %8, %3;"
               "mad.wide.u32 %0, %5, %8, %0; mad.wide.u32 %1, %6, %8, %1; mad.wide.u32 %2, %7, %8, %2; mad.wide.u32 %3, %4, %8, %3;"
               : "+l"(x0), "+l"(x1), "+l"(x2), "+l"(x3) : "r"(a), "r"(c), "r"(e), "r"(g), "r"(k));
  a = (uint32_t)(x0 >> 32); b = (uint32_t)x0; c = (uint32_t)(x1 >> 32); d = (uint32_t)x1;
  e = (uint32_t)(x2 >> 32); f = (uint32_t)x2; g = (uint32_t)(x3 >> 32); h = (uint32_t)x3;
})
BODY(k_iadd3, asm volatile("{ .reg .u32 t; add.u32 t, %0, %4; add.u32 %0, t, %8; add.u32 t, %1, %5; add.u32 %1, t, %8;"
                           "add.u32 t, %2, %6; add.u32 %2, t, %8; add.u32 t, %3, %7; add.u32 %3, t, %8;"
                           "add.u32 t, %4, %1; add.u32 %4, t, %8; add.u32 t, %5, %2; add.u32 %5, t, %8;"
                           "add.u32 t, %6, %3; add.u32 %6, t, %8; add.u32 t, %7, %0; add.u32 %7, t, %8; }"
                           : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k));)
BODY(k_lop3, asm volatile("lop3.b32 %0, %0, %4, %8, 0x96; lop3.b32 %1, %1, %5, %8, 0x96; lop3.b32 %2, %2, %6, %8, 0x96; lop3.b32 %3, %3, %7, %8, 0x96;"
                          "lop3.b32 %4, %4, %1, %8, 0xe8; lop3.b32 %5, %5, %2, %8, 0xe8; lop3.b32 %6, %6, %3, %8, 0xe8; lop3.b32 %7, %7, %0, %8, 0xe8;"
                          : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k));)
BODY(k_shf, asm volatile("shf.l.wrap.b32 %0, %0, %4, %8; shf.l.wrap.b32 %1, %1, %5, %8; shf.l.wrap.b32 %2, %2, %6, %8; shf.l.wrap.b32 %3, %3, %7, %8;"
                         "shf.l.wrap.b32 %4, %4, %1, %8; shf.l.wrap.b32 %5, %5, %2, %8; shf.l.wrap.b32 %6, %6, %3, %8; shf.l.wrap.b32 %7, %7, %0, %8;"
                         : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k));)
BODY(k_shl, asm volatile("shl.b32 %0, %0, %8; shl.b32 %1, %1, %8; shl.b32 %2, %2, %8; shl.b32 %3, %3, %8;"
                         "shr.u32 %4, %4, %8; shr.u32 %5, %5, %8; shr.u32 %6, %6, %8; shr.u32 %7, %7, %8;"
                         : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k & 1));)
BODY(k_flo, asm volatile("bfind.shiftamt.u32 %0, %4; bfind.shiftamt.u32 %1, %5; bfind.shiftamt.u32 %2, %6; bfind.shiftamt.u32 %3, %7;"
                         "bfind.shiftamt.u32 %4, %1; bfind.shiftamt.u32 %5, %2; bfind.shiftamt.u32 %6, %3; bfind.shiftamt.u32 %7, %0;"
                         : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h));)
BODY(k_sel, asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %8; selp.u32 %0, %1, %4, p; selp.u32 %1, %2, %5, p; selp.u32 %2, %3, %6, p;"
                         "selp.u32 %3, %4, %7, p; selp.u32 %4, %5, %0, p; selp.u32 %5, %6, %1, p; selp.u32 %6, %7, %2, p; }"
                         : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k));)
// half IMAD, half LOP3: do the two pipes overlap?
BODY(k_mix, asm volatile("mad.lo.u32 %0, %0, %8, %4; lop3.b32 %1, %1, %5, %8, 0x96; mad.lo.u32 %2, %2, %8, %6; lop3.b32 %3, %3, %7, %8, 0x96;"
                         "mad.lo.u32 %4, %4, %8, %0; lop3.b32 %5, %5, %1, %8, 0xe8; mad.lo.u32 %6, %6, %8, %2; lop3.b32 %7, %7, %3, %8, 0xe8;"
                         : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k));)
// one IMAD to three LOP3
BODY(k_mix13, asm volatile("mad.lo.u32 %0, %0, %8, %4; lop3.b32 %1, %1, %5, %8, 0x96; lop3.b32 %2, %2, %6, %8, 0x96; lop3.b32 %3, %3, %7, %8, 0x96;"
                           "mad.lo.u32 %4, %4, %8, %0; lop3.b32 %5, %5, %1, %8, 0xe8; lop3.b32 %6, %6, %2, %8, 0xe8; lop3.b32 %7, %7, %3, %8, 0xe8;"
                           : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k));)


BODY(k_imad_iadd, asm volatile("mad.lo.u32 %0, %0, %8, %4; add.u32 %1, %1, %5; mad.lo.u32 %2, %2, %8, %6; add.u32 %3, %3, %7;"
                         "mad.lo.u32 %4, %4, %8, %0; add.u32 %5, %5, %1; mad.lo.u32 %6, %6, %8, %2; add.u32 %7, %7, %3;"
                         : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k));)
BODY(k_imad_shf, asm volatile("mad.lo.u32 %0, %0, %8, %4; shf.l.wrap.b32 %1, %1, %5, %8; mad.lo.u32 %2, %2, %8, %6; shf.l.wrap.b32 %3, %3, %7, %8;"
                         "mad.lo.u32 %4, %4, %8, %0; shf.l.wrap.b32 %5, %5, %1, %8; mad.lo.u32 %6, %6, %8, %2; shf.l.wrap.b32 %7, %7, %3, %8;"
                         : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k));)
BODY(k_lop_shf, asm volatile("lop3.b32 %0, %0, %4, %8, 0x96; shf.l.wrap.b32 %1, %1, %5, %8; lop3.b32 %2, %2, %6, %8, 0x96; shf.l.wrap.b32 %3, %3, %7, %8;"
                         "lop3.b32 %4, %4, %0, %8, 0xe8; shf.l.wrap.b32 %5, %5, %1, %8; lop3.b32 %6, %6, %2, %8, 0xe8; shf.l.wrap.b32 %7, %7, %3, %8;"
                         : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k));)
BODY(k_ffma, { float fa = __uint_as_float(a), fb = __uint_as_float(b), fc = __uint_as_float(c), fd = __uint_as_float(d), fe = __uint_as_float(e), ff = __uint_as_float(f), fg = __uint_as_float(g), fh = __uint_as_float(h), fk = __uint_as_float(k);
  asm volatile("fma.rn.f32 %0, %0, %8, %4; fma.rn.f32 %1, %1, %8, %5; fma.rn.f32 %2, %2, %8, %6; fma.rn.f32 %3, %3, %8, %7;"
               "fma.rn.f32 %4, %4, %8, %0; fma.rn.f32 %5, %5, %8, %1; fma.rn.f32 %6, %6, %8, %2; fma.rn.f32 %7, %7, %8, %3;"
               : "+f"(fa), "+f"(fb), "+f"(fc), "+f"(fd), "+f"(fe), "+f"(ff), "+f"(fg), "+f"(fh) : "f"(fk));
  a = __float_as_uint(fa); b = __float_as_uint(fb); c = __float_as_uint(fc); d = __float_as_uint(fd); e = __float_as_uint(fe); f = __float_as_uint(ff); g = __float_as_uint(fg); h = __float_as_uint(fh); })
BODY(k_ffma_lop, { float fa = __uint_as_float(a), fc = __uint_as_float(c), fe = __uint_as_float(e), fg = __uint_as_float(g), fk = __uint_as_float(k);
  asm volatile("fma.rn.f32 %0, %0, %8, %2; lop3.b32 %4, %4, %5, %9, 0x96; fma.rn.f32 %1, %1, %8, %3; lop3.b32 %5, %5, %6, %9, 0x96;"
               "fma.rn.f32 %2, %2, %8, %0; lop3.b32 %6, %6, %7, %9, 0xe8; fma.rn.f32 %3, %3, %8, %1; lop3.b32 %7, %7, %4, %9, 0xe8;"
               : "+f"(fa), "+f"(fc), "+f"(fe), "+f"(fg), "+r"(b), "+r"(d), "+r"(f), "+r"(h) : "f"(fk), "r"(k));
  a = __float_as_uint(fa); c = __float_as_uint(fc); e = __float_as_uint(fe); g = __float_as_uint(fg); })
BODY(k_ffma_imad, { float fa = __uint_as_float(a), fc = __uint_as_float(c), fe = __uint_as_float(e), fg = __uint_as_float(g), fk = __uint_as_float(k);
  asm volatile("fma.rn.f32 %0, %0, %8, %2; mad.lo.u32 %4, %4, %9, %5; fma.rn.f32 %1, %1, %8, %3; mad.lo.u32 %5, %5, %9, %6;"
               "fma.rn.f32 %2, %2, %8, %0; mad.lo.u32 %6, %6, %9, %7; fma.rn.f32 %3, %3, %8, %1; mad.lo.u32 %7, %7, %9, %4;"
               : "+f"(fa), "+f"(fc), "+f"(fe), "+f"(fg), "+r"(b), "+r"(d), "+r"(f), "+r"(h) : "f"(fk), "r"(k));
  a = __float_as_uint(fa); c = __float_as_uint(fc); e = __float_as_uint(fe); g = __float_as_uint(fg); })
BODY(k_hi_lop, asm volatile("mul.hi.u32 %0, %0, %8; lop3.b32 %1, %1, %5, %8, 0x96; lop3.b32 %2, %2, %6, %8, 0x96; lop3.b32 %3, %3, %7, %8, 0x96;"
                         "mul.hi.u32 %4, %4, %8; lop3.b32 %5, %5, %1, %8, 0xe8; lop3.b32 %6, %6, %2, %8, 0xe8; lop3.b32 %7, %7, %3, %8, 0xe8;"
                         : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k));)
BODY(k_flo_lop, asm volatile("bfind.shiftamt.u32 %0, %4; lop3.b32 %1, %1, %5, %8, 0x96; lop3.b32 %2, %2, %6, %8, 0x96; lop3.b32 %3, %3, %7, %8, 0x96;"
                         "lop3.b32 %4, %4, %0, %8, 0x96; lop3.b32 %5, %5, %1, %8, 0xe8; lop3.b32 %6, %6, %2, %8, 0xe8; lop3.b32 %7, %7, %3, %8, 0xe8;"
                         : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k));)


// ---- round 2: which instructions issue at full rate?  (candidates for replacing the half-rate LOP3 / SHF / SEL / IMAD)
BODY(k_fsel, { float fa = __uint_as_float(a), fb = __uint_as_float(b), fc = __uint_as_float(c), fd = __uint_as_float(d), fe = __uint_as_float(e), ff = __uint_as_float(f), fg = __uint_as_float(g), fh = __uint_as_float(h);
  asm volatile("{ .reg .pred p; setp.lt.u32 p, %8, 77; selp.f32 %0, %1, %4, p; selp.f32 %1, %2, %5, p; selp.f32 %2, %3, %6, p; selp.f32 %3, %4, %7, p;"
               "selp.f32 %4, %5, %0, p; selp.f32 %5, %6, %1, p; selp.f32 %6, %7, %2, p; selp.f32 %7, %0, %3, p; }"
               : "+f"(fa), "+f"(fb), "+f"(fc), "+f"(fd), "+f"(fe), "+f"(ff), "+f"(fg), "+f"(fh) : "r"(k + i));
  a = __float_as_uint(fa); b = __float_as_uint(fb); c = __float_as_uint(fc); d = __float_as_uint(fd); e = __float_as_uint(fe); f = __float_as_uint(ff); g = __float_as_uint(fg); h = __float_as_uint(fh); })
BODY(k_selp, asm volatile("{ .reg .pred p; setp.lt.u32 p, %8, 77; selp.u32 %0, %1, %4, p; selp.u32 %1, %2, %5, p; selp.u32 %2, %3, %6, p; selp.u32 %3, %4, %7, p;"
                          "selp.u32 %4, %5, %0, p; selp.u32 %5, %6, %1, p; selp.u32 %6, %7, %2, p; selp.u32 %7, %0, %3, p; }"
                          : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k + i));)
BODY(k_fadd, { float fa = __uint_as_float(a), fb = __uint_as_float(b), fc = __uint_as_float(c), fd = __uint_as_float(d), fe = __uint_as_float(e), ff = __uint_as_float(f), fg = __uint_as_float(g), fh = __uint_as_float(h);
  asm volatile("add.rn.f32 %0, %0, %4; add.rn.f32 %1, %1, %5; add.rn.f32 %2, %2, %6; add.rn.f32 %3, %3, %7;"
               "mul.rn.f32 %4, %4, %0; mul.rn.f32 %5, %5, %1; mul.rn.f32 %6, %6, %2; mul.rn.f32 %7, %7, %3;"
               : "+f"(fa), "+f"(fb), "+f"(fc), "+f"(fd), "+f"(fe), "+f"(ff), "+f"(fg), "+f"(fh));
  a = __float_as_uint(fa); b = __float_as_uint(fb); c = __float_as_uint(fc); d = __float_as_uint(fd); e = __float_as_uint(fe); f = __float_as_uint(ff); g = __float_as_uint(fg); h = __float_as_uint(fh); })
BODY(k_lea, asm volatile("{ .reg .u32 t; shl.b32 t, %4, 16; add.u32 %0, %0, t; shl.b32 t, %5, 16; add.u32 %1, %1, t; shl.b32 t, %6, 16; add.u32 %2, %2, t; shl.b32 t, %7, 16; add.u32 %3, %3, t;"
                         "shl.b32 t, %0, 16; add.u32 %4, %4, t; shl.b32 t, %1, 16; add.u32 %5, %5, t; shl.b32 t, %2, 16; add.u32 %6, %6, t; shl.b32 t, %3, 16; add.u32 %7, %7, t; }"
                         : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h));)
BODY(k_prmt, asm volatile("prmt.b32 %0, %0, %4, %8; prmt.b32 %1, %1, %5, %8; prmt.b32 %2, %2, %6, %8; prmt.b32 %3, %3, %7, %8;"
                          "prmt.b32 %4, %4, %0, %8; prmt.b32 %5, %5, %1, %8; prmt.b32 %6, %6, %2, %8; prmt.b32 %7, %7, %3, %8;"
                          : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k & 0x7777));)
BODY(k_add_pred, asm volatile("{ .reg .pred p; setp.lt.u32 p, %8, 77; @p add.u32 %0, %0, %4; @p add.u32 %1, %1, %5; @p add.u32 %2, %2, %6; @p add.u32 %3, %3, %7;"
                              "@!p add.u32 %4, %4, %0; @!p add.u32 %5, %5, %1; @!p add.u32 %6, %6, %2; @!p add.u32 %7, %7, %3; }"
                              : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h) : "r"(k + i));)
BODY(k_setp, asm volatile("{ .reg .pred p0, p1, p2, p3; setp.lt.u32 p0, %0, %4; setp.lt.u32 p1, %1, %5; setp.lt.u32 p2, %2, %6; setp.lt.u32 p3, %3, %7;"
                          "@p0 add.u32 %4, %4, %0; @p1 add.u32 %5, %5, %1; @p2 add.u32 %6, %6, %2; @p3 add.u32 %7, %7, %3; }"
                          : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h));)
BODY(k_iadd3x, asm volatile("{ .reg .u32 t; add.cc.u32 %0, %0, %4; addc.u32 %1, %1, %5; add.cc.u32 %2, %2, %6; addc.u32 %3, %3, %7;"
                            "add.cc.u32 %4, %4, %1; addc.u32 %5, %5, %0; add.cc.u32 %6, %6, %3; addc.u32 %7, %7, %2; }"
                            : "+r"(a), "+r"(b), "+r"(c), "+r"(d), "+r"(e), "+r"(f), "+r"(g), "+r"(h));)
BODY(k_i2f, { float fa, fb, fc, fd;
  asm volatile("cvt.rn.f32.u32 %0, %4; cvt.rn.f32.u32 %1, %5; cvt.rn.f32.u32 %2, %6; cvt.rn.f32.u32 %3, %7;" : "=f"(fa), "=f"(fb), "=f"(fc), "=f"(fd) : "r"(a), "r"(b), "r"(c), "r"(d));
  asm volatile("cvt.rzi.u32.f32 %0, %4; cvt.rzi.u32.f32 %1, %5; cvt.rzi.u32.f32 %2, %6; cvt.rzi.u32.f32 %3, %7;" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "f"(fa), "f"(fb), "f"(fc), "f"(fd)); })
BODY(k_ffma_rm, { float fa = __uint_as_float(a), fb = __uint_as_float(b), fc = __uint_as_float(c), fd = __uint_as_float(d), fe = __uint_as_float(e), ff = __uint_as_float(f), fg = __uint_as_float(g), fh = __uint_as_float(h), fk = __uint_as_float(k);
  asm volatile("fma.rm.f32 %0, %0, %8, %4; fma.rm.f32 %1, %1, %8, %5; fma.rm.f32 %2, %2, %8, %6; fma.rm.f32 %3, %3, %8, %7;"
               "fma.rm.f32 %4, %4, %8, %0; fma.rm.f32 %5, %5, %8, %1; fma.rm.f32 %6, %6, %8, %2; fma.rm.f32 %7, %7, %8, %3;"
               : "+f"(fa), "+f"(fb), "+f"(fc), "+f"(fd), "+f"(fe), "+f"(ff), "+f"(fg), "+f"(fh) : "f"(fk));
  a = __float_as_uint(fa); b = __float_as_uint(fb); c = __float_as_uint(fc); d = __float_as_uint(fd); e = __float_as_uint(fe); f = __float_as_uint(ff); g = __float_as_uint(fg); h = __float_as_uint(fh); })
BODY(k_dfma, { double da = (double)a, db = (double)b, dc = (double)c, dd = (double)d, dk = 1.0000001;
  asm volatile("fma.rn.f64 %0, %0, %4, %1; fma.rn.f64 %1, %1, %4, %2; fma.rn.f64 %2, %2, %4, %3; fma.rn.f64 %3, %3, %4, %0;"
               "fma.rn.f64 %0, %0, %4, %2; fma.rn.f64 %1, %1, %4, %3; fma.rn.f64 %2, %2, %4, %0; fma.rn.f64 %3, %3, %4, %1;"
               : "+d"(da), "+d"(db), "+d"(dc), "+d"(dd) : "d"(dk));
  a = (uint32_t)da; b = (uint32_t)db; c = (uint32_t)dc; d = (uint32_t)dd; })
BODY(k_dfma_lop, { double da = (double)a, db = (double)b, dk = 1.0000001;
  asm volatile("fma.rn.f64 %0, %0, %6, %1; lop3.b32 %2, %2, %3, %7, 0x96; lop3.b32 %3, %3, %4, %7, 0x96; lop3.b32 %4, %4, %5, %7, 0xe8;"
               "fma.rn.f64 %1, %1, %6, %0; lop3.b32 %5, %5, %2, %7, 0x96; lop3.b32 %2, %2, %4, %7, 0x96; lop3.b32 %3, %3, %5, %7, 0xe8;"
               : "+d"(da), "+d"(db), "+r"(c), "+r"(d), "+r"(e), "+r"(f) : "d"(dk), "r"(k));
  a = (uint32_t)da; b = (uint32_t)db; })

typedef void (*kern_t)(uint32_t *, uint32_t, unsigned long long *);

int main()
{
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t *out;
  unsigned long long *cyc, h[1024];
  cudaMalloc(&out, sms * 1024 * 4);
  cudaMalloc(&cyc, sms * 8);
  struct { const char *name; kern_t k; int per_stmt; } tests[] = {
    { "IMAD (mad.lo)", k_imad, 8 }, { "IMAD.HI (mul.hi)", k_imadhi, 8 }, { "IMAD.WIDE (mad.wide)", k_imadwide, 8 }, { "IADD (add, 2 per chain)", k_iadd3, 16 },
    { "LOP3", k_lop3, 8 }, { "SHF (funnel)", k_shf, 8 }, { "SHL/SHR", k_shl, 8 }, { "FLO (bfind.shiftamt)", k_flo, 8 }, { "SETP+SELP", k_sel, 8 },
    { "IMAD+LOP3 1:1", k_mix, 8 }, { "IMAD+LOP3 1:3", k_mix13, 8 },
    { "IMAD+IADD 1:1", k_imad_iadd, 8 }, { "IMAD+SHF 1:1", k_imad_shf, 8 }, { "LOP3+SHF 1:1", k_lop_shf, 8 }, { "FFMA", k_ffma, 8 },
    { "FFMA+LOP3 1:1", k_ffma_lop, 8 }, { "FFMA+IMAD 1:1", k_ffma_imad, 8 }, { "IMAD.HI+LOP3 1:3", k_hi_lop, 8 }, { "FLO+LOP3 1:7", k_flo_lop, 8 },
    { "FSEL (selp.f32)", k_fsel, 8 }, { "SEL (selp.u32)", k_selp, 8 }, { "FADD/FMUL", k_fadd, 8 }, { "SHL 16 + ADD (LEA?)", k_lea, 8 }, { "PRMT", k_prmt, 8 },
    { "predicated IADD", k_add_pred, 8 }, { "ISETP + @p IADD", k_setp, 8 }, { "IADD.CC + IADDC", k_iadd3x, 8 }, { "I2F + F2I", k_i2f, 8 },
    { "FFMA.RM", k_ffma_rm, 8 }, { "DFMA", k_dfma, 8 }, { "DFMA+LOP3 1:3", k_dfma_lop, 8 },
  };
  for (auto &t : tests)
  {
    t.k<<<sms, 1024>>>(out, 12345, cyc);
    t.k<<<sms, 1024>>>(out, 12345, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; i++) avg += (double)h[i];
    avg /= sms;
    const double warp_instr_per_smsp = 8.0 * ITER * 4 * t.per_stmt;   // 8 warps per scheduler
    printf("%-28s %8.3f cycles per warp instruction per scheduler  (%s)\n", t.name, avg / warp_instr_per_smsp, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
