// Host harness for bench/fq_inv32.cuh (driven from Python: see the command in DESIGN / the commit message).
#include "../dusk-plonk_b200/csrc/fq_inv32.cuh"
extern "C" void fqi_inv(const uint32_t* y, const uint32_t* m, uint32_t mni, int rounds, uint32_t* out, int n32) {
    if (n32 == 8) fqinv::inv_mod<8>(y, m, mni, rounds, out);
    else fqinv::inv_mod<12>(y, m, mni, rounds, out);
}
