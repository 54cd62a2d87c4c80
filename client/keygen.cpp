// client/keygen.cpp -- `make keygen`: writes secret.key and eval.key in the working directory.
// Replaces client/gen_secure_keyset.cpp:94-120 (redsec_params_small_v2: n=350, N=1024, l=10, Bgbit=3, t=9, basebit=3).
// Randomness comes from the OS (ChaCha20 keyed by getrandom); `--seed N` selects the deterministic TEST generator instead
// (the reference seeds its generator with {0,0,0}, gen_secure_keyset.cpp:99: every run makes the same key).
// `-seclevel` is accepted and ignored, as the reference's program does (client/Makefile:3).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "redsec_b200.h"

int main(int argc, char** argv) {
    bool seeded = false;
    uint64_t seed = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--seed") && i + 1 < argc) { seeded = true; seed = strtoull(argv[++i], nullptr, 10); }
        else if (!strcmp(argv[i], "-seclevel") && i + 1 < argc) i++;
    }
    std::vector<int32_t> lwe_key(RS_LWE_N), tlwe_key(RS_TLWE_N);
    std::vector<uint32_t> bsk(RS_BSK_WORDS), ksk(RS_KSK_WORDS);
    const int rc = seeded ? rs_keygen(seed, lwe_key.data(), tlwe_key.data(), bsk.data(), ksk.data())
                          : rs_keygen_secure(lwe_key.data(), tlwe_key.data(), bsk.data(), ksk.data());
    if (rc != RS_OK) { fprintf(stderr, "key generation failed (%d)\n", rc); return 1; }
    if (rs_write_secret_key("secret.key", lwe_key.data(), tlwe_key.data()) != RS_OK ||
        rs_write_eval_key("eval.key", bsk.data(), ksk.data()) != RS_OK) {
        fprintf(stderr, "cannot write secret.key / eval.key\n");
        return 1;
    }
    printf("wrote secret.key and eval.key (%s)\n", seeded ? "deterministic TEST keyset" : "OS entropy");
    return 0;
}
