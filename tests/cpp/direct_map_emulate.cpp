// CPU emulation of k_direct_map (TEST INFRASTRUCTURE): the same direct_map.hpp, the same tiling of the source zone and
// the same "whole block leaves when every thread has its match" control flow, thread loops replaced by plain loops.
// Build: g++ -O2 -ffp-contract=off (multiregionfoam_b200/build.py).
#include "../../multiregionfoam_b200/csrc/direct_map.hpp"

#include <algorithm>

extern "C" int emu_direct_map_build(int32_t nTo, const double* to, int32_t nFrom, const double* from, double tol, int32_t* map)
{
    const int kTile = 256;
    const double fourTol2 = 4.0 * tol * tol;
    int unmatched = 0;
    for (int32_t base = 0; base < nTo; base += kTile) // one block
    {
        const int nThreads = std::min<int32_t>(kTile, nTo - base);
        int found[kTile];
        for (int t = 0; t < nThreads; t++) found[t] = -1;
        for (int32_t j0 = 0; j0 < nFrom; j0 += kTile)
        {
            const int n = std::min<int32_t>(kTile, nFrom - j0);
            bool all = true;
            for (int t = 0; t < nThreads; t++)
            {
                const int32_t i = base + t;
                if (found[t] < 0)
                    for (int k = 0; k < n; k++)
                        if (dmap::matches(to[3 * i], to[3 * i + 1], to[3 * i + 2], from[3 * (j0 + k)], from[3 * (j0 + k) + 1],
                                          from[3 * (j0 + k) + 2], tol, fourTol2))
                        {
                            found[t] = j0 + k;
                            break;
                        }
                all = all && found[t] >= 0;
            }
            if (all) break;
        }
        for (int t = 0; t < nThreads; t++)
        {
            map[base + t] = found[t];
            unmatched += found[t] < 0;
        }
    }
    return unmatched;
}
