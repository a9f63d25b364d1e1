// Exactness check of bench/fq52.cuh on the host: (a * b * 2^-416 mod p) against the product's 32-bit-limb
// Montgomery arithmetic (a * b * 2^-384 mod p), related by the factor 2^-32 mod p.
//   g++ -O2 -std=c++17 -frounding-math -mfma -o /tmp/fq52_test bench/fq52_host_test.cpp && /tmp/fq52_test
#include <fenv.h>
#include <stdio.h>
#include <stdlib.h>

#include "../dusk-plonk_b200/csrc/arith.cuh"
#include "fq52.cuh"

using zkp::fq_t;

// canonical integer (12 x u32) <-> 52-bit double limbs
static fq52::el to52(const fq_t& c) {
    fq52::el r;
    unsigned __int128 acc = 0; int bits_have = 0, limb = 0;
    for (int i = 0; i < 8; i++) {
        while (bits_have < 52 && limb < 12) { acc |= (unsigned __int128)c.l[limb++] << bits_have; bits_have += 32; }
        r.l[i] = (double)(uint64_t)(acc & fq52::MASK);
        acc >>= 52; bits_have -= 52; if (bits_have < 0) bits_have = 0;
    }
    return r;
}
static fq_t from52(const fq52::el& e) {
    fq_t r = fq_t::zero();
    unsigned __int128 acc = 0; int have = 0, out = 0;
    for (int i = 0; i < 8; i++) {
        acc |= (unsigned __int128)(uint64_t)e.l[i] << have; have += 52;
        while (have >= 32 && out < 12) { r.l[out++] = (uint32_t)acc; acc >>= 32; have -= 32; }
    }
    while (out < 12) { r.l[out++] = (uint32_t)acc; acc >>= 32; }
    return r;
}

int main() {
    fesetround(FE_TOWARDZERO);
    srand(7);
    // 2^-32 mod p in Montgomery form: from_mont-style trick: mont_mul(x, 1) = x * 2^-384; we need canonical
    // a*b*2^-416 = canonical(a*b*2^-384) * 2^-32.  Compute inv32 = (2^32)^-1 mod p via Fermat in Montgomery form.
    fq_t two32 = zkp::from_u64<zkp::FqParams>(1ull << 32);
    fq_t inv32 = zkp::inverse(two32);               // Montgomery form of 2^-32
    int bad = 0;
    for (int it = 0; it < 20000; it++) {
        fq_t a, b;
        for (int i = 0; i < 12; i++) { a.l[i] = (uint32_t)rand() ^ ((uint32_t)rand() << 16); b.l[i] = (uint32_t)rand() ^ ((uint32_t)rand() << 16); }
        a.l[11] &= 0x0fffffff; b.l[11] &= 0x0fffffff;   // below p
        if (it == 0) { a = fq_t::zero(); }
        if (it == 1) { for (int i = 0; i < 12; i++) a.l[i] = zkp::FqParams::p(i); a.l[0] -= 1; b = a; }  // p - 1
        if (it == 2) { a = fq_t::zero(); a.l[0] = 1; }
        // reference: canonical a*b*2^-416 mod p.  (a*b*2^-384) = mont_mul(a, b) on raw canonical inputs;
        // then times 2^-32: mont_mul(x, M(2^-32)) = x * 2^-32.
        fq_t ref = (a * b) * inv32;
        fq52::el got52 = fq52::mul(to52(a), to52(b));
        fq_t got = from52(got52);
        if (!(got == ref)) { if (bad < 5) printf("mismatch at %d\n", it); bad++; }
    }
    printf("fq52 host test: %s (%d mismatches of 20000)\n", bad ? "FAILED" : "ok", bad);
    return bad != 0;
}
