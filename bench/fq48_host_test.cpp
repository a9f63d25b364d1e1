// Exactness check of bench/fq48.cuh on the host against the product's 32-bit-limb Montgomery
// multiplication (arith.cuh): both compute a b 2^-384 mod p.
//   g++ -O2 -std=c++17 -frounding-math -mfma -o /tmp/fq48_test bench/fq48_host_test.cpp && /tmp/fq48_test
#include <fenv.h>
#include <stdio.h>
#include <stdlib.h>

#include "../dusk-plonk_b200/csrc/arith.cuh"
#include "fq48.cuh"

using zkp::fq_t;

static fq48::el to48(const fq_t& c) {
    fq48::el r;
    for (int i = 0; i < 8; i++) {   // 48 bits = one and a half 32-bit limbs
        const int bit = 48 * i, w = bit / 32, off = bit % 32;
        unsigned __int128 v = c.l[w];
        if (w + 1 < 12) v |= (unsigned __int128)c.l[w + 1] << 32;
        if (w + 2 < 12) v |= (unsigned __int128)c.l[w + 2] << 64;
        r.l[i] = (double)(uint64_t)((v >> off) & 0xffffffffffffull);
    }
    return r;
}
static fq_t from48(const fq48::el& e) {
    fq_t r = fq_t::zero();
    unsigned __int128 acc = 0; int have = 0, out = 0;
    for (int i = 0; i < 8; i++) {
        acc |= (unsigned __int128)(uint64_t)e.l[i] << have; have += 48;
        while (have >= 32 && out < 12) { r.l[out++] = (uint32_t)acc; acc >>= 32; have -= 32; }
    }
    return r;
}

int main() {
    fesetround(FE_TOWARDZERO);
    srand(11);
    int bad = 0;
    const int T = 200000;
    for (int it = 0; it < T; it++) {
        fq_t a, b;
        for (int i = 0; i < 12; i++) { a.l[i] = (uint32_t)rand() ^ ((uint32_t)rand() << 16); b.l[i] = (uint32_t)rand() ^ ((uint32_t)rand() << 16); }
        a.l[11] &= 0x0fffffff; b.l[11] &= 0x0fffffff;   // below p
        if (it == 0) a = fq_t::zero();
        if (it == 1 || it == 2) { for (int i = 0; i < 12; i++) a.l[i] = zkp::FqParams::p(i); a.l[0] -= 1; if (it == 1) b = a; }  // p - 1
        if (it == 3) { a = fq_t::zero(); a.l[0] = 1; }
        if (it >= 4 && it < 40) {   // limbs at their extremes: all-ones patterns below p
            for (int i = 0; i < 11; i++) { a.l[i] = 0xffffffffu; b.l[i] = (it & 1) ? 0xffffffffu : 0u; }
            a.l[11] = 0x1a0111e0u - it; b.l[11] = 0x1a0111e0u - 2 * it;
        }
        const fq_t ref = a * b;     // raw inputs: a b 2^-384 mod p
        const fq_t got = from48(fq48::mul(to48(a), to48(b)));
        if (!(got == ref)) { if (bad < 5) printf("mismatch at %d\n", it); bad++; }
    }
    printf("fq48 host test: %s (%d mismatches of %d)\n", bad ? "FAILED" : "ok", bad, T);
    return bad != 0;
}
