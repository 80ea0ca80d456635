/* TEST INFRASTRUCTURE: a plain-C caller of the drop-in boundary (include/ssw.h, libssw.so) -- no Python, no ctypes.
 *
 * Replays the two doc examples of the reference crate (/root/reference/src/lib.rs:22-41 "Embedding a watermark" and
 * :43-66 "Extracting and testing for a watermark") and the assertions of /root/reference/tests/single_simple.rs on the
 * reference's own fixture, entry point by entry point:
 *     Writer::new(img, WriteConfig::default()).mark(&[&mark]).into_rgb8()   -> ssw_writer_new_rgb8 / _embed / _result_rgb8
 *     Reader::base / Reader::derived / reader.extract                      -> ssw_reader_base_rgb8 / _derived_rgb8 / _extract
 *     Tester::new(&extracted).similarity(&mark).exceeds_sigma(6.0)         -> ssw_similarity
 *     MarkBuf::generate_normal(1000)                                       -> ssw_mark_generate_normal
 *
 * usage: cabi_flow <cat.rgb8> <width> <height> <mark.f32> <n> <golden.rgb8> [<random_mark.f32>]
 * Inputs are raw dumps of tests/golden/*.npz written by tests/test_cabi_flow.py.  Exit code 0 = all checks passed;
 * every check prints one line. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ssw.h"

#define CHECK(call)                                                                          \
    do {                                                                                     \
        int rc_ = (call);                                                                    \
        if (rc_ != SSW_OK) {                                                                 \
            fprintf(stderr, "FAILED %s -> %d: %s\n", #call, rc_, ssw_last_error());          \
            return 2;                                                                        \
        }                                                                                    \
    } while (0)

static void* slurp(const char* path, size_t bytes) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(3); }
    void* p = malloc(bytes);
    if (fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read on %s\n", path); exit(3); }
    fclose(f);
    return p;
}

int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: %s cat.rgb8 w h mark.f32 n golden.rgb8 [random.f32]\n", argv[0]); return 3; }
    const uint32_t w = (uint32_t)atoi(argv[2]), h = (uint32_t)atoi(argv[3]);
    const size_t n = (size_t)atol(argv[5]), npix = (size_t)w * h;
    uint8_t* cat = (uint8_t*)slurp(argv[1], npix * 3);
    float* mark = (float*)slurp(argv[4], n * sizeof(float));
    uint8_t* golden = (uint8_t*)slurp(argv[6], npix * 3);
    int failures = 0;

    printf("%s\n", ssw_version());
    ssw_ctx* ctx = NULL;
    CHECK(ssw_ctx_create(0, &ctx));
    const ssw_config cfg = {SSW_METHOD_OPTION2, 0.1f, SSW_ORDER_ENERGY};   /* WriteConfig::default() / ReadConfig::default() */

    /* ---- lib.rs:22-41 / single_simple.rs:23-43: embed, compare with tests/watermarked_with_1.png */
    ssw_writer* wr = NULL;
    CHECK(ssw_writer_new_rgb8(ctx, cat, w, h, &cfg, &wr));
    const float* marks[1] = {mark};
    const size_t lens[1] = {n};
    CHECK(ssw_writer_embed(wr, marks, lens, 1));
    uint8_t* marked = (uint8_t*)malloc(npix * 3);
    CHECK(ssw_writer_result_rgb8(wr, marked));
    CHECK(ssw_writer_destroy(wr));
    size_t differ = 0;
    int maxd = 0;
    for (size_t i = 0; i < npix * 3; ++i) {
        const int d = abs((int)marked[i] - (int)golden[i]);
        differ += d != 0;
        if (d > maxd) maxd = d;
    }
    printf("embed vs golden PNG: %zu of %zu values differ, max |d| %d\n", differ, npix * 3, maxd);
    if (maxd > 1 || differ > 64) { printf("  FAIL: more than 64 +-1 LSB flips\n"); ++failures; }

    /* ---- lib.rs:43-66 / single_simple.rs:48-90: extract from the GOLDEN image, score */
    ssw_reader *base = NULL, *derived = NULL;
    CHECK(ssw_reader_base_rgb8(ctx, cat, w, h, &cfg, &base));
    CHECK(ssw_reader_derived_rgb8(ctx, golden, w, h, &derived));
    float* extracted = (float*)calloc(n, sizeof(float));
    CHECK(ssw_reader_extract(base, derived, extracted, n));
    float max_err = 0.f, sum_err = 0.f;
    for (size_t i = 0; i < n; ++i) {
        const float e = fabsf(extracted[i] - mark[i]);
        if (e > max_err) max_err = e;
        sum_err += e;
    }
    printf("extract: max abs error %.4f (< 0.12), mean abs error %.4f (< 0.02)\n", max_err, sum_err / (float)n);
    if (!(max_err < 0.12f) || !(sum_err / (float)n < 0.02f)) { printf("  FAIL\n"); ++failures; }
    float sim = 0.f;
    CHECK(ssw_similarity(ctx, extracted, mark, n, &sim));
    printf("similarity of the embedded mark: %.4f (> 31.2), exceeds 6 sigma: %s\n", sim, sim > 6.0f ? "yes" : "no");
    if (!(sim > 31.2f)) { printf("  FAIL\n"); ++failures; }
    if (argc > 7) {
        float* rnd = (float*)slurp(argv[7], n * sizeof(float));
        float rsim = 0.f;
        CHECK(ssw_similarity(ctx, extracted, rnd, n, &rsim));
        printf("similarity of an unrelated mark: %.4f (< 2.0)\n", rsim);
        if (!(rsim < 2.0f)) { printf("  FAIL\n"); ++failures; }
        free(rnd);
    }

    /* ---- the reference's misuse panics come back as status codes (src/algorithm.rs:530, :553-555) */
    if (ssw_reader_extract(derived, base, extracted, n) != SSW_ERR_STATE) { printf("FAIL: extract on a derived reader must be SSW_ERR_STATE\n"); ++failures; }
    if (ssw_reader_extract(base, derived, extracted, npix) != SSW_ERR_INVALID) { printf("FAIL: n >= w*h must be SSW_ERR_INVALID\n"); ++failures; }
    CHECK(ssw_reader_destroy(derived));
    CHECK(ssw_reader_destroy(base));

    /* ---- MarkBuf::generate_normal(1000) (src/algorithm.rs:619-626): a fresh mark embeds and is detected */
    float* fresh = (float*)malloc(n * sizeof(float));
    CHECK(ssw_mark_generate_normal(ctx, 0 /* entropy, like thread_rng */, n, fresh));
    CHECK(ssw_writer_new_rgb8(ctx, cat, w, h, &cfg, &wr));
    const float* fmarks[1] = {fresh};
    CHECK(ssw_writer_embed(wr, fmarks, lens, 1));
    CHECK(ssw_writer_result_rgb8(wr, marked));
    CHECK(ssw_writer_destroy(wr));
    CHECK(ssw_reader_base_rgb8(ctx, cat, w, h, &cfg, &base));
    CHECK(ssw_reader_derived_rgb8(ctx, marked, w, h, &derived));
    CHECK(ssw_reader_extract(base, derived, extracted, n));
    CHECK(ssw_similarity(ctx, extracted, fresh, n, &sim));
    printf("generate_normal mark: similarity %.4f (> 25)\n", sim);
    if (!(sim > 25.0f)) { printf("  FAIL\n"); ++failures; }
    CHECK(ssw_reader_destroy(derived));
    CHECK(ssw_reader_destroy(base));
    CHECK(ssw_ctx_destroy(ctx));
    free(cat); free(mark); free(golden); free(marked); free(extracted); free(fresh);
    printf(failures ? "cabi_flow: %d check(s) FAILED\n" : "cabi_flow: all checks passed\n", failures);
    return failures ? 1 : 0;
}
