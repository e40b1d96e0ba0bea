// CPU check of BallScale (bpldenoising_b200/csrc/common.cuh): the straight-line chain that replaces the IEEE sqrt and
// division of the projection `α / sqrt(n²)` in the strict kernels, run with a software stand-in for the hardware's
// reciprocal-square-root seed that has only the accuracy the hardware guarantees (emu_rsqrt_seed, emu_cuda.h), against
// the correctly rounded operations of the host.  Counts the operand pairs whose bits differ (expected: none).
#include "emu_cuda.h"
#include "../../bpldenoising_b200/csrc/common.cuh"

using namespace bpltv;

static unsigned long long mix(unsigned long long x)
{
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
static double mk64(int e, unsigned long long sig)
{
    unsigned long long b = ((unsigned long long)(1023 + e) << 52) | (sig & ((1ull << 52) - 1));
    double v; std::memcpy(&v, &b, 8); return v;
}
static float mk32(int e, unsigned long long sig)
{
    unsigned b = ((unsigned)(127 + e) << 23) | (unsigned)(sig & ((1u << 23) - 1));
    float v; std::memcpy(&v, &b, 4); return v;
}
static unsigned long long hard_sig(unsigned long long h, int S)
{
    const unsigned long long ones = (1ull << S) - 1, r = mix(h);
    switch (h & 7) {
    case 0: return ones - ((h >> 3) & 15);
    case 1: return (h >> 3) & 15;
    case 2: return 1ull << (r % S);
    case 3: return ones ^ (1ull << (r % S));
    case 4: return r & (ones << (S / 2));
    case 5: return r & (ones >> (S / 2));
    default: return r;
    }
}

// mode 0: image range, 1: whole range, 2: structured significands, 3 (fp32): every `a` pattern in turn.
// out[0] = pairs the chain took, out[1] = mismatches, out[2] / out[3] = first mismatching operands (as doubles)
extern "C" int emu_ballscale(int prec, int mode, unsigned long long count, unsigned long long seed, double *out)
{
    unsigned long long took = 0, bad = 0;
    out[2] = out[3] = 0.0;
    for (unsigned long long i = 0; i < count; ++i) {
        const unsigned long long h1 = mix(seed + 2 * i), h2 = mix(seed + 2 * i + 1);
        if (prec == 64) {
            double a, al;
            if (mode == 0) { a = mk64(-24 + (int)((h1 >> 54) % 30), h1); al = mk64(-14 + (int)((h2 >> 54) % 16), h2); }
            else if (mode == 1) { a = mk64(-499 + (int)((h1 >> 53) % 999), h1); al = mk64(-199 + (int)((h2 >> 53) % 399), h2); }
            else { a = mk64(-499 + (int)((h1 >> 53) % 999), hard_sig(h1, 52)); al = mk64(-199 + (int)((h2 >> 53) % 399), ((h2 >> 41) & 1) ? hard_sig(h2, 52) : h2); }
            if (!BallScale<double>::fast_ok(a, al)) continue;
            ++took;
            const double got = BallScale<double>::eval(a, al), want = al / std::sqrt(a);
            if (std::memcmp(&got, &want, 8) != 0 && bad++ == 0) { out[2] = a; out[3] = al; }
        } else {
            float a, al;
            if (mode == 0) { a = mk32(-24 + (int)((h1 >> 54) % 30), h1); al = mk32(-14 + (int)((h2 >> 54) % 16), h2); }
            else if (mode == 1) { a = mk32(-59 + (int)((h1 >> 53) % 119), h1); al = mk32(-29 + (int)((h2 >> 53) % 59), h2); }
            else if (mode == 2) { a = mk32(-59 + (int)((h1 >> 53) % 119), hard_sig(h1, 23)); al = mk32(-29 + (int)((h2 >> 53) % 59), ((h2 >> 41) & 1) ? hard_sig(h2, 23) : h2); }
            else { const unsigned long long k = i % (119ull << 23); a = mk32(-59 + (int)(k >> 23), k); al = mk32(-29 + (int)((h2 >> 53) % 59), h2); }
            if (!BallScale<float>::fast_ok(a, al)) continue;
            ++took;
            const float got = BallScale<float>::eval(a, al), want = al / std::sqrt(a);
            if (std::memcmp(&got, &want, 4) != 0 && bad++ == 0) { out[2] = a; out[3] = al; }
        }
    }
    out[0] = (double)took; out[1] = (double)bad;
    return 0;
}
