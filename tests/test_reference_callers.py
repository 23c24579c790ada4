"""CPU test (-m "not gpu"): the reference's own callers compile -- and link -- unchanged against this repository's
headers and host library.  SURVEY.md 8b names them: demo/main.cpp, test/render_test.cpp, test/scene/scene_test.cpp,
benchmark/main.cpp (plus the remaining unit tests and the two fuzz targets).  GoogleTest and Google Benchmark are not in
the image, so the test supplies two tiny stand-in headers that define the macros those files use; nothing is executed
(there is no GPU here), the point is that every declaration they rely on exists with a compatible signature and every
symbol they reference is exported by libPathTrace.so.  Skipped where the reference tree is not mounted (the GPU box)."""
import os
import subprocess

import pytest

from cpupathtrace_b200 import REPO_ROOT, lib_path

REFERENCE = "/root/reference"

GTEST_STUB = r"""
#pragma once
#include <cmath>
#include <iostream>
#define TEST(suite, name) void suite##_##name##_body()
struct StubStream { template<typename T> StubStream &operator<<(const T &) { return *this; } };
#define STUB_CHECK(cond) if(cond) {} else StubStream()
#define EXPECT_TRUE(x) STUB_CHECK(static_cast<bool>(x))
#define EXPECT_FALSE(x) STUB_CHECK(!static_cast<bool>(x))
#define ASSERT_TRUE(x) STUB_CHECK(static_cast<bool>(x))
#define ASSERT_FALSE(x) STUB_CHECK(!static_cast<bool>(x))
#define EXPECT_EQ(a, b) STUB_CHECK((a) == (b))
#define ASSERT_EQ(a, b) STUB_CHECK((a) == (b))
#define EXPECT_NE(a, b) STUB_CHECK((a) != (b))
#define ASSERT_NE(a, b) STUB_CHECK((a) != (b))
#define EXPECT_LT(a, b) STUB_CHECK((a) < (b))
#define EXPECT_LE(a, b) STUB_CHECK((a) <= (b))
#define EXPECT_GT(a, b) STUB_CHECK((a) > (b))
#define EXPECT_GE(a, b) STUB_CHECK((a) >= (b))
#define ASSERT_GT(a, b) STUB_CHECK((a) > (b))
#define ASSERT_LT(a, b) STUB_CHECK((a) < (b))
#define EXPECT_FLOAT_EQ(a, b) STUB_CHECK(std::fabs((a) - (b)) < 1e-6)
#define ASSERT_FLOAT_EQ(a, b) STUB_CHECK(std::fabs((a) - (b)) < 1e-6)
#define EXPECT_NEAR(a, b, tol) STUB_CHECK(std::fabs((a) - (b)) <= (tol))
#define ASSERT_NEAR(a, b, tol) STUB_CHECK(std::fabs((a) - (b)) <= (tol))
#define EXPECT_NO_THROW(stmt) stmt
#define EXPECT_ANY_THROW(stmt) try { stmt; } catch(...) {}
#define EXPECT_THROW(stmt, type) try { stmt; } catch(const type &) {}
namespace testing { inline void InitGoogleTest(int *, char **) {} }
inline int RUN_ALL_TESTS() { return 0; }
"""

GMOCK_STUB = r"""
#pragma once
#include <gtest/gtest.h>
#include <cmath>
namespace testing {
    template<typename V> struct EqMatcher { V v; template<typename T> bool operator()(const T &x) const { return x == v; } };
    template<typename V> struct GeMatcher { V v; template<typename T> bool operator()(const T &x) const { return x >= v; } };
    template<typename V> struct GtMatcher { V v; template<typename T> bool operator()(const T &x) const { return x > v; } };
    template<typename V> struct LtMatcher { V v; template<typename T> bool operator()(const T &x) const { return x < v; } };
    struct FloatEqMatcher { float v; bool operator()(float x) const { return std::fabs(x - v) <= 4e-7F * std::fabs(v); } };
    struct FloatNearMatcher { float v, tol; bool operator()(float x) const { return std::fabs(x - v) <= tol; } };
    struct NotNullMatcher { template<typename T> bool operator()(const T &x) const { return x != nullptr; } };
    template<typename V> EqMatcher<V> Eq(V v) { return {v}; }
    template<typename V> GeMatcher<V> Ge(V v) { return {v}; }
    template<typename V> GtMatcher<V> Gt(V v) { return {v}; }
    template<typename V> LtMatcher<V> Lt(V v) { return {v}; }
    inline FloatEqMatcher FloatEq(float v) { return {v}; }
    inline FloatNearMatcher FloatNear(float v, float tol) { return {v, tol}; }
    inline NotNullMatcher NotNull() { return {}; }
}
#define EXPECT_THAT(value, matcher) STUB_CHECK((matcher)(value))
#define ASSERT_THAT(value, matcher) STUB_CHECK((matcher)(value))
"""

BENCHMARK_STUB = r"""
#pragma once
#include <cstdint>
namespace benchmark {
    struct State {
        struct Iterator { bool operator!=(const Iterator &) const { return false; } void operator++() {} int operator*() const { return 0; } };
        Iterator begin() { return {}; }
        Iterator end() { return {}; }
        int64_t range(int) const { return 1; }
        void SetItemsProcessed(int64_t) {}
        int64_t iterations() const { return 1; }
    };
    enum TimeUnit { kNanosecond, kMicrosecond, kMillisecond, kSecond };
    struct Registration {
        Registration *Unit(TimeUnit) { return this; }
        Registration *Arg(int64_t) { return this; }
        Registration *Args(std::initializer_list<int64_t>) { return this; }
        Registration *UseRealTime() { return this; }
        Registration *Iterations(int64_t) { return this; }
        Registration *MinTime(double) { return this; }
    };
    template<typename T> void DoNotOptimize(T &&) {}
    inline void ClobberMemory() {}
    inline Registration *RegisterBenchmark(const char *, void (*)(State &)) { static Registration r; return &r; }
    inline void Initialize(int *, char **) {}
    inline bool ReportUnrecognizedArguments(int, char **) { return false; }
    inline int RunSpecifiedBenchmarks() { return 0; }
}
#define BENCHMARK(fn) static benchmark::Registration *registration_##fn = (new benchmark::Registration())
#define BENCHMARK_MAIN() int main() { return 0; }
"""


@pytest.fixture(scope="module")
def stubs(tmp_path_factory):
    if not os.path.isdir(os.path.join(REFERENCE, "demo")):
        pytest.skip("/root/reference is not mounted")
    root = tmp_path_factory.mktemp("stubs")
    (root / "gtest").mkdir()
    (root / "gtest" / "gtest.h").write_text(GTEST_STUB)
    (root / "gmock").mkdir()
    (root / "gmock" / "gmock.h").write_text(GMOCK_STUB)
    (root / "benchmark").mkdir()
    (root / "benchmark" / "benchmark.h").write_text(BENCHMARK_STUB)
    return root


def _compile(sources, stubs, out, extra=()):
    libdir = os.path.dirname(lib_path("libPathTrace.so"))
    cmd = ["g++", "-std=gnu++20", "-O0", "-w", f"-I{os.path.join(REPO_ROOT, 'include')}", f"-I{stubs}", f"-I{os.path.join(REFERENCE, 'test')}", *extra, *sources, "-o", str(out),
           f"-L{libdir}", "-lPathTrace", "-lptb", f"-Wl,-rpath,{libdir}", "-pthread"]
    done = subprocess.run(cmd, capture_output=True, text=True)
    assert done.returncode == 0, done.stderr[-3000:]


def test_demo_compiles_and_links_unchanged(stubs, tmp_path):
    _compile([os.path.join(REFERENCE, "demo", "main.cpp")], stubs, tmp_path / "demo")


def test_reference_unit_tests_compile_and_link_unchanged(stubs, tmp_path):
    tests = ["render_test.cpp", "post_processing_test.cpp", "scene/scene_test.cpp", "scene/boundig_box_test.cpp", "scene/mesh_test.cpp", "image/image_io_test.cpp",
             "test_utils.cpp"]
    main = tmp_path / "main.cpp"
    main.write_text("int main() { return 0; }\n")
    _compile([os.path.join(REFERENCE, "test", t) for t in tests] + [str(main)], stubs, tmp_path / "tests")


def test_benchmark_and_fuzz_targets_compile_and_link_unchanged(stubs, tmp_path):
    _compile([os.path.join(REFERENCE, "benchmark", "main.cpp")], stubs, tmp_path / "benchmark")
    for target in ("target_image_io_read.cpp", "target_mesh_parser.cpp"):
        main = tmp_path / f"main_{target}"
        main.write_text("#include <cstddef>\n#include <cstdint>\nextern \"C\" int LLVMFuzzerTestOneInput(const uint8_t *, size_t);\nint main() { return LLVMFuzzerTestOneInput(nullptr, 0); }\n")
        _compile([os.path.join(REFERENCE, "fuzz", target), "-x", "c++", str(main)], stubs, tmp_path / target.replace(".cpp", ""))
