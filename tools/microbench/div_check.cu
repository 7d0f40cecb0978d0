// Validation: lambda = a/b computed as q = a*r; e = fma(-b, q, a); fma(e, r, q) with
// r = RN(1/b) (Markstein's correction step) must equal the IEEE quotient bit for bit
// whenever the exponent guards of node_solve.cuh pass. Counts mismatches over random and
// adversarial operands.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30))*0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27))*0x94D049BB133111EBull; return z ^ (z >> 31);
}
__device__ __forceinline__ double make(uint64_t bits, int mode, int emin, int erange) {
  // sign | exponent in [emin, emin+erange) | mantissa (mode-dependent)
  uint64_t man = bits & 0xFFFFFFFFFFFFFull;
  if (mode == 1) man |= 0xFFFFFFFFF0000ull;          // mantissa near all ones
  if (mode == 2) man &= 0xFFull;                     // near power of two (low bits only)
  if (mode == 3) man = (man & 0xF) | 0xFFFFFFFFFFFF0ull;
  uint64_t ex = (uint64_t)(emin + (int)((bits >> 52) % erange));
  uint64_t sg = bits >> 63;
  return __longlong_as_double((long long)((sg << 63) | (ex << 52) | man));
}
__global__ void check(unsigned long long seed, int emin, int erange, unsigned long long* bad,
                      unsigned long long* guarded, int iters) {
  uint64_t s = mix(seed + blockIdx.x*(uint64_t)blockDim.x + threadIdx.x);
  unsigned long long nb = 0, ng = 0;
  for (int i = 0; i < iters; ++i) {
    s = mix(s); const uint64_t ba = s; s = mix(s); const uint64_t bb = s;
    const int mode = (int)((ba >> 60) & 3), modeb = (int)((bb >> 58) & 3);
    const double a = make(ba, mode == 3 ? 0 : mode, emin, erange), b = make(bb, modeb, 1023 - 40, 80);
    const double r = 1.0/b;
    const double q = a*r;
    const double e = fma(-b, q, a);
    const double l = fma(e, r, q);
    const unsigned ea = ((unsigned)__double2hiint(a) >> 20) & 0x7ff;
    const unsigned el = ((unsigned)__double2hiint(l) >> 20) & 0x7ff;
    const bool safe = (ea - 128u) < 1792u && (el - 128u) < 1792u;
    if (!safe) { ++ng; continue; }
    const double t = a/b;
    if (__double_as_longlong(t) != __double_as_longlong(l)) ++nb;
  }
  atomicAdd(bad, nb); atomicAdd(guarded, ng);
}
int main() {
  unsigned long long* d; cudaMalloc(&d, 16);
  const int iters = 4096;
  struct { int emin, erange; const char* what; } cases[] = {
    {1023 - 30, 60, "a exponent in [-30, 30)"}, {1, 2046, "a exponent over the whole range"},
    {1023 - 2, 4, "a, b same magnitude"}, {100, 60, "a tiny (guard region)"}};
  for (auto& c : cases) {
    cudaMemset(d, 0, 16);
    for (int rep = 0; rep < 8; ++rep)
      check<<<148*16, 256>>>(0x1234567ull*(rep + 1) + c.emin, c.emin, c.erange, d, d + 1, iters);
    unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-36s samples %.3e mismatches %llu guarded %llu\n", c.what, 8.0*148*16*256*iters, h[0], h[1]);
  }
  return 0;
}
