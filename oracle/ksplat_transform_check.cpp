// oracle/ksplat_transform_check.cpp -- TEST INFRASTRUCTURE ONLY.  Prints the band-1 / band-2 SH weight matrices that
// gaussiansplats3d_b200/csrc/ksplat_transform.h derives from a column-major 4x4 transform given on the command line (16 doubles),
// one value per line, so a test can compare them with the Python restatement (scenes.sh_rotation_matrices).
#include <cstdio>
#include <cstdlib>
#include "../gaussiansplats3d_b200/csrc/ksplat_transform.h"

int main(int argc, char **argv) {
    if (argc != 17) return 2;
    double e[16];
    for (int i = 0; i < 16; ++i) e[i] = atof(argv[1 + i]);
    gs::KTransform K;
    gs::ksplat_transform_params(e, -1.5, 1.5, K);
    for (int l = 0; l < 3; ++l) for (int k = 0; k < 3; ++k) printf("%.17g\n", K.m1[l][k]);
    for (int l = 0; l < 5; ++l) for (int k = 0; k < 5; ++k) printf("%.17g\n", K.m2[l][k]);
    return 0;
}
