// tests/ref_dropin/gtest/gtest.h — TEST INFRASTRUCTURE.  A few-line stand-in for googletest (not in this image, and the
// reference fetches it from the network: reference tests/CMakeLists.txt:5-12) so that the reference's own test files
// compile, unmodified and from where they lie, against include/polympc_compat/.  Single translation unit per test binary:
// main() lives here.  No `inline` keyword in this file on purpose (see polympc_compat.hpp, functor annotation).
#pragma once
#include <cmath>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing {
struct Case { std::string name; void (*body)(); };
struct State {
    std::vector<Case> cases;
    int failures = 0;
    static State& get() { static State s; return s; }
};
struct Registrar { Registrar(const char* name, void (*body)()) { State::get().cases.push_back(Case{name, body}); } };
/** result of one expectation; `<< "message"` is accepted like in googletest */
struct Check {
    bool ok;
    std::ostringstream msg;
    Check(bool good, const char* file, int line, const char* text) : ok(good)
    { if (!ok) { State::get().failures++; msg << file << ":" << line << ": Failure\n  " << text; } }
    Check(const Check& o) : ok(o.ok) { msg << o.msg.str(); }
    ~Check() { if (!ok) std::cout << msg.str() << std::endl; }
    template <class T> Check& operator<<(const T& v) { if (!ok) msg << " " << v; return *this; }
};
struct Test {};
static int RunAll()
{
    State& s = State::get();
    for (const Case& c : s.cases) {
        const int before = s.failures;
        std::cout << "[ RUN      ] " << c.name << std::endl;
        c.body();
        std::cout << (s.failures == before ? "[       OK ] " : "[  FAILED  ] ") << c.name << std::endl;
    }
    std::cout << "[==========] " << s.cases.size() << " tests, " << s.failures << " failed expectations" << std::endl;
    return s.failures == 0 ? 0 : 1;
}
static void InitGoogleTest(int*, char**) {}
} // namespace testing

#define TEST(suite, name)                                                                             \
    static void suite##_##name##_Body();                                                              \
    static ::testing::Registrar suite##_##name##_registrar(#suite "." #name, &suite##_##name##_Body); \
    static void suite##_##name##_Body()
#define GTEST_CHECK_(cond, text) ::testing::Check((cond), __FILE__, __LINE__, text)
#define EXPECT_TRUE(c) GTEST_CHECK_(static_cast<bool>(c), "Expected true: " #c)
#define EXPECT_FALSE(c) GTEST_CHECK_(!static_cast<bool>(c), "Expected false: " #c)
#define EXPECT_EQ(a, b) GTEST_CHECK_((a) == (b), "Expected equality: " #a " == " #b)
#define EXPECT_NE(a, b) GTEST_CHECK_((a) != (b), "Expected: " #a " != " #b)
#define EXPECT_LT(a, b) GTEST_CHECK_((a) < (b), "Expected: " #a " < " #b)
#define EXPECT_LE(a, b) GTEST_CHECK_((a) <= (b), "Expected: " #a " <= " #b)
#define EXPECT_GT(a, b) GTEST_CHECK_((a) > (b), "Expected: " #a " > " #b)
#define EXPECT_GE(a, b) GTEST_CHECK_((a) >= (b), "Expected: " #a " >= " #b)
#define EXPECT_NEAR(a, b, tol) GTEST_CHECK_(std::fabs((a) - (b)) <= (tol), "Expected: |" #a " - " #b "| <= " #tol)
#define ASSERT_TRUE EXPECT_TRUE
#define ASSERT_FALSE EXPECT_FALSE
#define ASSERT_EQ EXPECT_EQ
#define ASSERT_LT EXPECT_LT
#define ASSERT_LE EXPECT_LE
#define ASSERT_NEAR EXPECT_NEAR
#define RUN_ALL_TESTS() ::testing::RunAll()

#ifndef GTEST_SHIM_NO_MAIN
int main(int argc, char** argv) { ::testing::InitGoogleTest(&argc, argv); return RUN_ALL_TESTS(); }
#endif
