// oracle/ref_harness.cpp -- OUR driver around the reference's UNMODIFIED plaintext build (lib/*.cpp + nets/<net>/net.cpp,
// compiled from /root/reference where they lie by oracle/build_ref.sh; outputs only under oracle/_ref/).
// Test infrastructure: prints the raw class scores of HeBNN::run for CSV rows so the golden vectors in
// tests/golden/ can be (re)generated.  Mirrors the input handling of nets/mnist/sign1024x1/main.cpp:159-176
// (x = 2*pixel - 255) or, with a third argument "relu", of nets/mnist/relu1024x1/main.cpp:191-207 (x = pixel/100 - 1).
// Usage: ptxt_<net> <csv> [max_rows] [relu]   (run with cwd = the net directory: it opens var_prep.dat)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "Layer.h"
#include "net.h"

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s <csv> [max_rows]\n", argv[0]); return 2; }
    const int max_rows = argc > 2 ? atoi(argv[2]) : 1;
    const bool relu_input = argc > 3 && !strcmp(argv[3], "relu");
    HeBNN* network = new HeBNN();
    tDimensions indim, outdim;
    network->get_in_dims(&indim);
    network->get_out_dims(&outdim);
    const size_t len = (size_t)indim.hw.h * indim.hw.w * indim.in_dep;
    FILE* fd = fopen(argv[1], "r");
    if (!fd) { fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    std::vector<char> line(len * 5 + 64);
    int rows = 0;
    while (rows < max_rows && fgets(line.data(), (int)line.size(), fd)) {
        if (line[0] < '0' || line[0] > '9') continue;
        tFixedPoint* data = (tFixedPoint*)calloc(len, sizeof(tFixedPoint));
        char* tok = strtok(line.data(), ",");
        const int label = atoi(tok);
        for (size_t i = 0; i < len; i++) {
            tok = strtok(NULL, ",\n");
            if (tok && *tok) data[i] = relu_input ? (tFixedPoint)((int)(atoi(tok) / 100) - 1) : (tFixedPoint)(2 * atoi(tok) - 255);
        }
        tFixedPoint* res = (tFixedPoint*)network->run(data);
        printf("label %d scores", label);
        for (uint32_t i = 0; i < outdim.in_dep; i++) printf(" %d", (int)res[i]);
        printf("\n");
        free(res);
        rows++;
    }
    fclose(fd);
    return 0;
}
