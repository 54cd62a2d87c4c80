// client/decrypt_image.cpp -- `make decrypt-image`: network_output.ctxt + secret.key -> class scores and
// "Classification Result: d".  Replaces client/decrypt_image.cpp:46-63 (lweSymDecrypt + modSwitchFromTorus32 with message
// space 4096, centred to (-2048, 2048], argmax).  Argument: MNIST | CIFAR-10 | ImageNet (10 / 10 / 1000 classes).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <strings.h>
#include <vector>

#include "redsec_b200.h"

int main(int argc, char** argv) {
    int classes = 10;
    if (argc > 1 && !strcasecmp(argv[1], "imagenet")) classes = 1000;
    else if (argc > 1 && strcasecmp(argv[1], "mnist") && strcasecmp(argv[1], "cifar-10")) { printf("Invalid data format!\n"); return 1; }
    std::vector<int32_t> lwe_key(RS_LWE_N), tlwe_key(RS_TLWE_N);
    if (rs_read_secret_key("secret.key", lwe_key.data(), tlwe_key.data()) != RS_OK) { fprintf(stderr, "cannot read secret.key\n"); return 1; }
    std::vector<uint32_t> ct((size_t)classes * RS_LWE_WORDS);
    if (rs_read_ctxt("network_output.ctxt", ct.data(), classes) != RS_OK) { fprintf(stderr, "cannot read network_output.ctxt\n"); return 1; }
    std::vector<int32_t> scores(classes);
    if (rs_lwe_decrypt(scores.data(), ct.data(), classes, lwe_key.data(), 4096) != RS_OK) return 1;
    printf("Scores:");
    for (int v : scores) printf(" %d", v);
    printf("\nClassification Result: %d\n", (int)(std::max_element(scores.begin(), scores.end()) - scores.begin()));
    return 0;
}
