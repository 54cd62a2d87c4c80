// client/encrypt_image.cpp -- `make encrypt-image`: image.ptxt (label,h,w,c,p0,p1,..., as client/image_converter.py writes it)
// + secret.key -> image.ctxt, one LWE sample per pixel.
// Replaces client/encrypt_image.cpp:65-85: ptxt = 2*pixel - 255, mu = modSwitchToTorus32(ptxt, 4096), alpha = 2^-15.
// ALL h*w*c pixels are encrypted: the reference's loop only consumes comma-terminated fields and leaves the last pixel
// unencrypted; that defect is not reproduced.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "redsec_b200.h"

int main(int argc, char** argv) {
    const char* path = "image.ptxt";
    bool seeded = false;
    uint64_t seed = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--seed") && i + 1 < argc) { seeded = true; seed = strtoull(argv[++i], nullptr, 10); }
        else path = argv[i];
    }
    std::ifstream in(path);
    if (!in) { fprintf(stderr, "cannot open %s\n", path); return 1; }
    std::stringstream ss;
    ss << in.rdbuf();
    std::vector<long> fields;
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        if (tok.find_first_of("0123456789") == std::string::npos) continue;
        fields.push_back(strtol(tok.c_str(), nullptr, 10));
    }
    if (fields.size() < 4) { fprintf(stderr, "%s: expected label,h,w,c,pixels...\n", path); return 1; }
    const size_t count = (size_t)fields[1] * fields[2] * fields[3];
    if (fields.size() != 4 + count) { fprintf(stderr, "%s: %zu pixels, header says %zu\n", path, fields.size() - 4, count); return 1; }
    std::vector<int32_t> lwe_key(RS_LWE_N), tlwe_key(RS_TLWE_N);
    if (rs_read_secret_key("secret.key", lwe_key.data(), tlwe_key.data()) != RS_OK) { fprintf(stderr, "cannot read secret.key\n"); return 1; }
    std::vector<uint32_t> mu(count), ct(count * RS_LWE_WORDS);
    for (size_t i = 0; i < count; i++) mu[i] = rs_modswitch_to_torus32((int32_t)(fields[4 + i] * 2 - 255), 4096);
    const double alpha = std::ldexp(1.0, -15);      // SECALPHA, client/encrypt_image.cpp:10
    const int rc = seeded ? rs_lwe_encrypt(ct.data(), mu.data(), count, alpha, lwe_key.data(), seed)
                          : rs_lwe_encrypt_secure(ct.data(), mu.data(), count, alpha, lwe_key.data());
    if (rc != RS_OK || rs_write_ctxt("image.ctxt", ct.data(), count, alpha * alpha, 0) != RS_OK) { fprintf(stderr, "encryption failed\n"); return 1; }
    printf("encrypted %zu pixels (label %ld) into image.ctxt\n", count, fields[0]);
    return 0;
}
