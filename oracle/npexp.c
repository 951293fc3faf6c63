/*
 * oracle/npexp.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * CPU restatement of the float64 `exp` that numpy 2.3.5 executes on AVX512_SKX hosts: numpy's
 * DOUBLE_exp_AVX512_SKX loop calls Intel SVML `__svml_exp8_ha` (vendored into numpy's
 * _multiarray_umath; a third-party dependency of the reference, not under /root/reference).
 * The reference reaches it through `1 - np.exp(-spatial_distance_map)` and
 * `1 - np.exp(-temporal_distance_map)` (det3d/selectors/spatial_temporal_selector.py:141-143).
 *
 * Algorithm of the main path (|x| < 0x1.61da04cbafe44p+9), read from the shipped binary:
 *   z  = fma_rz(x, log2e, Shifter)           Shifter = 1.5*2^48 (+bias bits): keeps 4 fraction bits,
 *                                            round-toward-zero of a positive sum == floor
 *   N  = z - Shifter                         N = floor(16 x log2e)/16 = m + j/16
 *   r  = fma(-N, ln2_hi, x); r = fma(-ln2_lo, N, r)          r in [0, ln2/16)
 *   P  = ((c6 r + c5) r^2 + (c4 r + c3)) r^2 + (c2 r + c1)    (Estrin, all fma)
 *   y  = fma(Th[j], fma(P, r, Tl[j]), Th[j])                  Th+Tl ~ 2^(j/16)
 *   exp(x) = scalef(y, N) = y * 2^floor(N)
 * Every constant below was read out of the binary's __svml_dexp_ha_data_internal_avx512 table.
 * The restatement is pinned bit-for-bit against np.exp by tests/test_oracle_npexp.py.
 * |x| >= threshold (SVML's scalar "rare" path: overflow/underflow/denormal results) is NOT pinned
 * bitwise: it falls back to libm exp.  The selector never depends on those bits: for any
 * e < 2^-54, 1 - e rounds to exactly 1.0.
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fPIC -shared (contraction must stay off: the
 * fused/unfused split is part of the function's definition).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static const uint64_t TH[16] = {
    0x3ff0000000000000ull, 0x3ff0b5586cf9890full, 0x3ff172b83c7d517bull, 0x3ff2387a6e756238ull,
    0x3ff306fe0a31b715ull, 0x3ff3dea64c123422ull, 0x3ff4bfdad5362a27ull, 0x3ff5ab07dd485429ull,
    0x3ff6a09e667f3bcdull, 0x3ff7a11473eb0187ull, 0x3ff8ace5422aa0dbull, 0x3ff9c49182a3f090ull,
    0x3ffae89f995ad3adull, 0x3ffc199bdd85529cull, 0x3ffd5818dcfba487ull, 0x3ffea4afa2a490daull,
};
static const uint64_t TL[16] = {
    0x0000000000000000ull, 0x3c979aa65d837b6dull, 0xbc801b15eaa59348ull, 0x3c968efde3a8a894ull,
    0x3c834d754db0abb6ull, 0x3c859f48a72a4c6dull, 0x3c7690cebb7aafb0ull, 0x3c9063e1e21c5409ull,
    0xbc93b3efbf5e2228ull, 0xbc7b32dcb94da51dull, 0x3c8db72fc1f0eab4ull, 0x3c71affc2b91ce27ull,
    0x3c8c1a7792cb3387ull, 0x3c736eae30af0cb3ull, 0x3c74a385a63d07a7ull, 0xbc8ff7128fd391f0ull,
};

static inline double u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

double oracle_npexp(double x)
{
    const double L2E = 0x1.71547652b82fep+0, SHIFTER = 0x1.8000000003ff0p+48;
    const double L2H = 0x1.62e42fefa39efp-1, L2L = 0x1.abc9e3b39803fp-56;
    const double C6 = 0x1.7411836940c04p-10, C5 = 0x1.1101cbbc265c0p-7, C4 = 0x1.55557242d68fep-5,
                 C3 = 0x1.5555553939732p-3, C2 = 0x1.000000000d008p-1, C1 = 0x1.fffffffffff70p-1;
    if (!(fabs(x) < 0x1.61da04cbafe44p+9)) return exp(x);    /* rare path: not pinned (see header) */

    /* z = round-toward-zero(x*L2E + SHIFTER).  The exact sum is positive, so RZ == floor onto z's
     * grid (multiples of 2^-4).  Emulated in round-to-nearest with an exact error term. */
    double hi = x * L2E;
    double lo = fma(x, L2E, -hi);                /* hi + lo == x*L2E exactly */
    double z = SHIFTER + hi;                     /* RN onto the 2^-4 grid */
    double err = (hi - (z - SHIFTER)) + lo;      /* exact sign of (x*L2E + SHIFTER) - z */
    if (err < 0.0) z -= 0.0625;
    double N = z - SHIFTER;
    unsigned j = (unsigned)(d2u(z) & 0xf);
    double r = fma(-N, L2H, x);
    r = fma(-L2L, N, r);
    double r2 = r * r;
    double p65 = fma(C6, r, C5);
    double p43 = fma(C4, r, C3);
    double p21 = fma(C2, r, C1);
    double P = fma(r2, p65, p43);
    P = fma(r2, P, p21);
    double th = u2d(TH[j]), tl = u2d(TL[j]);
    double y = fma(P, r, tl);
    y = fma(th, y, th);
    return ldexp(y, (int)floor(N));              /* vscalefpd(y, N) */
}

void oracle_npexp_array(const double *x, double *out, long n)
{
    for (long i = 0; i < n; ++i) out[i] = oracle_npexp(x[i]);
}

/* 1 - exp(-d): the reference's "exp" normalisation (spatial_temporal_selector.py:141-143). */
void oracle_exp_normalize_array(const double *d, double *out, long n)
{
    for (long i = 0; i < n; ++i) out[i] = 1.0 - oracle_npexp(-d[i]);
}

/*
 * float32 `exp` of numpy 2.3.5 on AVX512F / AVX2+FMA3 hosts: numpy's own SIMD kernel (simd_exp_FLOAT in
 * loops_exponent_log.dispatch.c.src; float32 does not go through SVML).  The reference reaches it through
 * `1 - np.exp(-feature_distance_map)` on the float32 map torch hands over
 * (det3d/selectors/spatial_feature_selector.py:190, spatial_temporal_feature_selector.py:214).
 *
 *   q   = rint(x * log2e)                      (add/subtract 1.5*2^23)
 *   r   = fma(q, -ln2_hi, x); r = fma(q, -ln2_lo, r)          Cody-Waite, ln2_hi = 6.93145752e-1f
 *   num = ((((p5 r + p4) r + p3) r + p2) r + p1) r + p0        all fma
 *   den = (q2 r + q1) r + q0
 *   exp = scalef(num / den, q)                 x > 88.7228... -> inf, x < -103.972... -> 0, NaN -> NaN
 *
 * Pinned bit for bit against np.exp on ALL 2^32 float32 inputs in the build container (NaN payloads
 * aside; tests/test_oracle_npexp.py re-checks a sample and the edges, ADSEL_EXHAUSTIVE=1 the whole range).
 */
float oracle_npexpf(float x)
{
    const float xmax = 88.72283935546875f, xmin = -103.97208404541015625f;
    const float log2e = 1.44269504088896341f;
    const float c1 = -6.93145752e-1f, c2 = -1.42860677e-6f;
    const float p0 = 9.999999999980870924916e-01f, p1 = 7.257664613233124478488e-01f,
                p2 = 2.473615434895520810817e-01f, p3 = 5.114512081637298353406e-02f,
                p4 = 6.757896990527504603057e-03f, p5 = 5.082762527590693718096e-04f;
    const float q0 = 1.0f, q1 = -2.742335390411667452936e-01f, q2 = 2.159509375685829852307e-02f;
    const float magic = 0x1.8p+23f;
    if (x != x) return x;
    if (x > xmax) return INFINITY;
    if (x < xmin) return 0.0f;
    float q = x * log2e;
    q = (q + magic) - magic;
    float r = fmaf(q, c1, x);
    r = fmaf(q, c2, r);
    float num = fmaf(p5, r, p4);
    num = fmaf(num, r, p3);
    num = fmaf(num, r, p2);
    num = fmaf(num, r, p1);
    num = fmaf(num, r, p0);
    float den = fmaf(q2, r, q1);
    den = fmaf(den, r, q0);
    return ldexpf(num / den, (int)q);            /* vscalefps: one rounding, denormal results included */
}

void oracle_npexpf_array(const float *x, float *out, long n)
{
    for (long i = 0; i < n; ++i) out[i] = oracle_npexpf(x[i]);
}

/* 1 - exp(-d) in float32: the feature term's normalisation. */
void oracle_exp_normalize_f32_array(const float *d, float *out, long n)
{
    for (long i = 0; i < n; ++i) out[i] = 1.0f - oracle_npexpf(-d[i]);
}
