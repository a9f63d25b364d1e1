// Host build of arith.cuh / g1.cuh (PTX carry flag emulated) so the exact device algorithms
// can be unit-tested without a GPU.  Built and driven by tests/test_host_arith.py.
#include "arith.cuh"
#include "g1.cuh"
#include "host_inv.h"
#include <cstring>
using namespace zkp;

template <class F> static F load(const uint32_t* p) { F r; std::memcpy(r.l, p, sizeof(r.l)); return r; }
template <class F> static void store(uint32_t* p, const F& v) { std::memcpy(p, v.l, sizeof(v.l)); }

extern "C" {
void ht_fr_mul(const uint32_t* a, const uint32_t* b, uint32_t* o) { store(o, load<fr_t>(a) * load<fr_t>(b)); }
void ht_fr_add(const uint32_t* a, const uint32_t* b, uint32_t* o) { store(o, load<fr_t>(a) + load<fr_t>(b)); }
void ht_fr_sub(const uint32_t* a, const uint32_t* b, uint32_t* o) { store(o, load<fr_t>(a) - load<fr_t>(b)); }
void ht_fr_neg(const uint32_t* a, uint32_t* o) { store(o, neg(load<fr_t>(a))); }
void ht_fr_inv(const uint32_t* a, uint32_t* o) { store(o, inverse(load<fr_t>(a))); }
void ht_fr_from_mont(const uint32_t* a, uint32_t* o) { store(o, from_mont(load<fr_t>(a))); }
void ht_fr_to_mont(const uint32_t* a, uint32_t* o) { store(o, to_mont(load<fr_t>(a))); }
void ht_fq_mul(const uint32_t* a, const uint32_t* b, uint32_t* o) { store(o, load<fq_t>(a) * load<fq_t>(b)); }
void ht_fq_add(const uint32_t* a, const uint32_t* b, uint32_t* o) { store(o, load<fq_t>(a) + load<fq_t>(b)); }
void ht_fq_sub(const uint32_t* a, const uint32_t* b, uint32_t* o) { store(o, load<fq_t>(a) - load<fq_t>(b)); }
void ht_fq_neg(const uint32_t* a, uint32_t* o) { store(o, neg(load<fq_t>(a))); }
void ht_fq_inv(const uint32_t* a, uint32_t* o) { store(o, inverse(load<fq_t>(a))); }
void ht_fq_to_mont(const uint32_t* a, uint32_t* o) { store(o, to_mont(load<fq_t>(a))); }

// XYZZ accumulate: acc (x,y,zz,zzz : 4 x 12 limbs) += affine (x,y : 2 x 12 limbs), neg flag
void ht_xyzz_madd(uint32_t* acc, const uint32_t* aff, int negate) {
    g1_xyzz A; std::memcpy(&A, acc, sizeof(A));
    g1_affine Q; std::memcpy(&Q, aff, sizeof(Q));
    if (negate) Q.y = neg(Q.y);
    xyzz_madd(A, Q);
    std::memcpy(acc, &A, sizeof(A));
}
void ht_xyzz_add(uint32_t* acc, const uint32_t* other) {
    g1_xyzz A, B; std::memcpy(&A, acc, sizeof(A)); std::memcpy(&B, other, sizeof(B));
    xyzz_add(A, B);
    std::memcpy(acc, &A, sizeof(A));
}
void ht_xyzz_dbl(uint32_t* acc) {
    g1_xyzz A; std::memcpy(&A, acc, sizeof(A));
    xyzz_dbl(A);
    std::memcpy(acc, &A, sizeof(A));
}
void ht_xyzz_to_affine(const uint32_t* acc, uint32_t* aff) {
    g1_xyzz A; std::memcpy(&A, acc, sizeof(A));
    g1_affine Q = xyzz_to_affine(A);
    std::memcpy(aff, &Q, sizeof(Q));
}
// plain modular inverse by the binary-GCD routine of host_inv.h: y, m, out as N64 x u64 (N64 = 4 or 6)
void ht_inv_mod(const uint64_t* y, const uint64_t* m, uint64_t* out, int n64) {
    if (n64 == 4) hostinv::inv_mod<4>(y, m, out);
    else hostinv::inv_mod<6>(y, m, out);
}
}
