/* cmocka.h -- TEST INFRASTRUCTURE ONLY: a minimal stand-in for the cmocka API subset that the
 * reference's test/*.c programs use (assert_int_equal, assert_true, assert_non_null,
 * assert_null, assert_string_equal, assert_ptr_equal, assert_ptr_not_equal, cmocka_unit_test,
 * cmocka_run_group_tests, struct CMUnitTest).  cmocka itself is not installed in this image;
 * with this header the reference's unit tests compile UNMODIFIED against this library
 * (tests/test_reference_c_tests.py).  Written from the public cmocka API names only. */
#ifndef HUF_TEST_CMOCKA_SHIM_H
#define HUF_TEST_CMOCKA_SHIM_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef void (*CMUnitTestFunction)(void **state);

struct CMUnitTest {
    const char *name;
    CMUnitTestFunction test_func;
};

#define cmocka_unit_test(f) { #f, f }

static int shim_failures;

#define SHIM_FAIL(...)                                            \
    do {                                                          \
        fprintf(stderr, "%s:%d: ", __FILE__, __LINE__);           \
        fprintf(stderr, __VA_ARGS__);                             \
        fprintf(stderr, "\n");                                    \
        shim_failures++;                                          \
    } while (0)

#define assert_true(c) do { if (!(c)) SHIM_FAIL("assert_true(%s)", #c); } while (0)
#define assert_int_equal(a, b)                                                             \
    do {                                                                                   \
        long long a__ = (long long)(a), b__ = (long long)(b);                              \
        if (a__ != b__) SHIM_FAIL("assert_int_equal(%s, %s): %lld != %lld", #a, #b, a__, b__); \
    } while (0)
#define assert_non_null(p) do { if ((p) == NULL) SHIM_FAIL("assert_non_null(%s)", #p); } while (0)
#define assert_null(p) do { if ((p) != NULL) SHIM_FAIL("assert_null(%s)", #p); } while (0)
#define assert_string_equal(a, b)                                                          \
    do {                                                                                   \
        if (strcmp((const char *)(a), (const char *)(b)) != 0)                             \
            SHIM_FAIL("assert_string_equal(%s, %s)", #a, #b);                              \
    } while (0)
#define assert_ptr_equal(a, b) do { if ((const void *)(a) != (const void *)(b)) SHIM_FAIL("assert_ptr_equal(%s, %s)", #a, #b); } while (0)
#define assert_ptr_not_equal(a, b) do { if ((const void *)(a) == (const void *)(b)) SHIM_FAIL("assert_ptr_not_equal(%s, %s)", #a, #b); } while (0)

static int shim_run_group(const struct CMUnitTest *tests, size_t n)
{
    size_t i;
    for (i = 0; i < n; i++) {
        void *state = NULL;
        int before = shim_failures;
        tests[i].test_func(&state);
        printf("[ %s ] %s\n", shim_failures == before ? "  OK  " : "FAILED", tests[i].name);
    }
    return shim_failures ? 1 : 0;
}

#define cmocka_run_group_tests(tests, setup, teardown) \
    shim_run_group((tests), sizeof(tests) / sizeof((tests)[0]))

#endif
