// CPU check of median13.cuh (host build of the __host__ __device__ programs).
//   nvcc -O2 -o /tmp/test_median13 tools/test_median13.cu && /tmp/test_median13
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../katsdpsigproc_b200/csrc/median13.cuh"

int main()
{
    srand(1);
    long bad = 0;
    for (int trial = 0; trial < 300000; trial++) {
        float e[16];
        int mod = (trial % 3 == 0) ? 5 : 100000;  // many ties in a third of the trials
        for (int k = 0; k < 16; k++) e[k] = (float) (rand() % mod) * 0.37f;
        int rot = trial & 15;
        float r[16];
        for (int k = 0; k < 16; k++) r[(k + rot) & 15] = e[k];
        float o[4];
        ksp::median13x4(r, rot, o[0], o[1], o[2], o[3]);
        for (int j = 0; j < 4; j++) {
            std::vector<float> w(e + j, e + j + 13);
            std::sort(w.begin(), w.end());
            if (w[6] != o[j]) bad++;
        }
        {   // eight medians from 20 samples
            float e20[20], o8[8];
            for (int k = 0; k < 20; k++) e20[k] = (float) (rand() % mod) * 0.37f;
            ksp::median13x8(e20, o8);
            for (int j = 0; j < 8; j++) {
                std::vector<float> w(e20 + j, e20 + j + 13);
                std::sort(w.begin(), w.end());
                if (w[6] != o8[j]) bad++;
            }
        }
        float w13[13];
        for (int k = 0; k < 13; k++) w13[k] = e[k];
        unsigned valid = (unsigned) rand() & 0x1fff;
        float lo, hi;
        bool ok = ksp::median_masked13(w13, valid, lo, hi);
        std::vector<float> v;
        for (int k = 0; k < 13; k++) if ((valid >> k) & 1) v.push_back(e[k]);
        std::sort(v.begin(), v.end());
        if (ok != !v.empty()) bad++;
        if (ok && (lo != v[(v.size() - 1) / 2] || hi != v[v.size() / 2])) bad++;
    }
    printf("mismatches: %ld\n", bad);
    return bad != 0;
}
