// gtest/gtest.h -- a minimal stand-in for GoogleTest, test infrastructure only.
//
// The reference's conformance suite (/root/reference/tests/*.cpp|*.cu, 48 cases) is written
// against gtest 1.14, which the reference fetches over the network (CMakeLists.txt:32-39) and
// which this image neither has nor can download.  This header implements exactly the subset
// those files use -- TEST, TEST_F, ::testing::Test (SetUp/TearDown), EXPECT_/ASSERT_ {EQ, NE, LT,
// LE, GT, GE, TRUE, FALSE, FLOAT_EQ, NEAR, STREQ, NO_THROW} with streamed messages, and a main()
// with --gtest_filter / --gtest_list_tests -- so that the reference's test sources compile
// UNCHANGED against libspmv_b200.so (tests/ref_suite/Makefile, tests/test_reference_suite.py).
// Semantics follow gtest: EXPECT_* records a failure and continues, ASSERT_* returns from the
// test body, EXPECT_FLOAT_EQ accepts 4 ULPs, a thrown exception fails the test.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <functional>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

namespace testing {

class Message {
public:
    template <class T>
    Message& operator<<(const T& v) { ss_ << v; return *this; }
    Message& operator<<(std::ostream& (*f)(std::ostream&)) { ss_ << f; return *this; }
    std::string str() const { return ss_.str(); }
private:
    std::ostringstream ss_;
};

namespace internal {

struct State {
    int failures_in_current = 0;
    static State& get() { static State s; return s; }
};

struct Result {
    bool ok;
    std::string text;
    explicit operator bool() const { return ok; }
};

// operator= is what makes `EXPECT_x(...) << "msg"` work: the streamed Message binds tighter.
class Reporter {
public:
    Reporter(const char* file, int line, const std::string& text) : file_(file), line_(line), text_(text) {}
    void operator=(const Message& m) const {
        ++State::get().failures_in_current;
        std::string extra = m.str();
        std::printf("%s:%d: Failure\n%s%s%s\n", file_, line_, text_.c_str(), extra.empty() ? "" : "\n", extra.c_str());
        std::fflush(stdout);
    }
private:
    const char* file_;
    int line_;
    std::string text_;
};

template <class T, class = void>
struct Streamable : std::false_type {};
template <class T>
struct Streamable<T, std::void_t<decltype(std::declval<std::ostream&>() << std::declval<const T&>())>> : std::true_type {};

template <class T>
std::string show(const T& v) {
    std::ostringstream ss;
    if constexpr (std::is_same_v<std::decay_t<T>, std::nullptr_t>) {
        ss << "nullptr";
    } else if constexpr (std::is_same_v<std::decay_t<T>, bool>) {
        ss << (v ? "true" : "false");
    } else if constexpr (std::is_floating_point_v<std::decay_t<T>>) {
        ss.precision(std::numeric_limits<std::decay_t<T>>::max_digits10);
        ss << v;
    } else if constexpr (std::is_enum_v<std::decay_t<T>>) {
        ss << static_cast<long long>(static_cast<std::underlying_type_t<std::decay_t<T>>>(v));
    } else if constexpr (std::is_pointer_v<std::decay_t<T>> && !std::is_same_v<std::decay_t<T>, const char*> &&
                         !std::is_same_v<std::decay_t<T>, char*>) {
        ss << static_cast<const void*>(v);
    } else if constexpr (Streamable<T>::value) {
        ss << v;
    } else {
        ss << "<" << sizeof(T) << "-byte object>";
    }
    return ss.str();
}

template <class A, class B, class Op>
Result compare(const char* ea, const char* eb, const A& a, const B& b, const char* opname, Op op) {
    if (op(a, b)) return {true, ""};
    std::ostringstream ss;
    ss << "Expected: (" << ea << ") " << opname << " (" << eb << "), actual: " << show(a) << " vs " << show(b);
    return {false, ss.str()};
}

#define GTEST_SHIM_CMP_(name, opsym)                                                              \
    template <class A, class B>                                                                   \
    Result name(const char* ea, const char* eb, const A& a, const B& b) {                         \
        return compare(ea, eb, a, b, #opsym, [](const A& x, const B& y) { return x opsym y; });   \
    }
GTEST_SHIM_CMP_(cmp_eq, ==)
GTEST_SHIM_CMP_(cmp_ne, !=)
GTEST_SHIM_CMP_(cmp_lt, <)
GTEST_SHIM_CMP_(cmp_le, <=)
GTEST_SHIM_CMP_(cmp_gt, >)
GTEST_SHIM_CMP_(cmp_ge, >=)
#undef GTEST_SHIM_CMP_

inline Result check_bool(const char* e, bool v, bool want) {
    if (v == want) return {true, ""};
    std::ostringstream ss;
    ss << "Value of: " << e << "\n  Actual: " << (v ? "true" : "false") << "\nExpected: " << (want ? "true" : "false");
    return {false, ss.str()};
}

// gtest's AlmostEquals: sign-magnitude -> biased integers, at most 4 units in the last place apart
inline bool almost_equal_4ulp(float a, float b) {
    if (std::isnan(a) || std::isnan(b)) return false;
    auto biased = [](float f) {
        uint32_t u;
        std::memcpy(&u, &f, 4);
        return (u & 0x80000000u) ? (~u + 1u) : (u | 0x80000000u);
    };
    const uint32_t x = biased(a), y = biased(b);
    return (x > y ? x - y : y - x) <= 4u;
}

inline Result cmp_float_eq(const char* ea, const char* eb, float a, float b) {
    if (almost_equal_4ulp(a, b)) return {true, ""};
    std::ostringstream ss;
    ss << "Expected equality of these values:\n  " << ea << "\n    Which is: " << show(a) << "\n  " << eb
       << "\n    Which is: " << show(b);
    return {false, ss.str()};
}

inline Result cmp_near(const char* ea, const char* eb, const char* et, double a, double b, double tol) {
    const double d = std::fabs(a - b);
    if (d <= tol) return {true, ""};
    std::ostringstream ss;
    ss << "The difference between " << ea << " and " << eb << " is " << d << ", which exceeds " << et << ", where\n"
       << ea << " evaluates to " << show(a) << ",\n" << eb << " evaluates to " << show(b) << ", and\n" << et
       << " evaluates to " << show(tol) << ".";
    return {false, ss.str()};
}

inline Result cmp_streq(const char* ea, const char* eb, const char* a, const char* b) {
    const bool same = (a == nullptr || b == nullptr) ? a == b : std::strcmp(a, b) == 0;
    if (same) return {true, ""};
    std::ostringstream ss;
    ss << "Expected equality of these values:\n  " << ea << "\n    Which is: " << (a ? a : "NULL") << "\n  " << eb
       << "\n    Which is: " << (b ? b : "NULL");
    return {false, ss.str()};
}

struct TestInfo {
    std::string suite, name;
    std::function<void()> run;
};
inline std::vector<TestInfo>& registry() { static std::vector<TestInfo> r; return r; }
struct Registrar {
    Registrar(const char* suite, const char* name, std::function<void()> run) { registry().push_back({suite, name, std::move(run)}); }
};

// '*' / '?' glob
inline bool glob(const char* p, const char* s) {
    if (*p == '\0') return *s == '\0';
    if (*p == '*') return glob(p + 1, s) || (*s != '\0' && glob(p, s + 1));
    if (*s == '\0') return false;
    return (*p == '?' || *p == *s) && glob(p + 1, s + 1);
}
inline bool any_pattern(const std::string& list, const std::string& name) {
    size_t i = 0;
    while (i <= list.size()) {
        size_t j = list.find(':', i);
        if (j == std::string::npos) j = list.size();
        if (j > i && glob(list.substr(i, j - i).c_str(), name.c_str())) return true;
        i = j + 1;
    }
    return false;
}
// positive[:positive...][-negative[:negative...]]
inline bool selected(const std::string& filter, const std::string& name) {
    std::string pos = filter, neg;
    const size_t dash = filter.find('-');
    if (dash != std::string::npos) { pos = filter.substr(0, dash); neg = filter.substr(dash + 1); }
    if (pos.empty()) pos = "*";
    return any_pattern(pos, name) && !(neg.size() && any_pattern(neg, name));
}

inline std::string& filter_flag() { static std::string f = "*"; return f; }
inline bool& list_flag() { static bool b = false; return b; }

}  // namespace internal

class Test {
public:
    virtual ~Test() {}
    virtual void SetUp() {}
    virtual void TearDown() {}
    virtual void TestBody() = 0;
    void Run() {
        SetUp();
        if (internal::State::get().failures_in_current == 0) TestBody();
        TearDown();
    }
};

inline void InitGoogleTest(int* argc, char** argv) {
    for (int i = 1; i < *argc; ++i) {
        const std::string a = argv[i];
        if (a.rfind("--gtest_filter=", 0) == 0) internal::filter_flag() = a.substr(15);
        else if (a == "--gtest_list_tests") internal::list_flag() = true;
    }
}

inline int RunAllTests() {
    using namespace internal;
    int ran = 0, failed = 0;
    std::vector<std::string> failed_names;
    for (auto& t : registry()) {
        const std::string full = t.suite + "." + t.name;
        if (!selected(filter_flag(), full)) continue;
        if (list_flag()) { std::printf("%s\n", full.c_str()); continue; }
        std::printf("[ RUN      ] %s\n", full.c_str());
        std::fflush(stdout);
        State::get().failures_in_current = 0;
        try {
            t.run();
        } catch (const std::exception& e) {
            ++State::get().failures_in_current;
            std::printf("unknown file: Failure\nC++ exception with description \"%s\" thrown in the test body.\n", e.what());
        } catch (...) {
            ++State::get().failures_in_current;
            std::printf("unknown file: Failure\nUnknown C++ exception thrown in the test body.\n");
        }
        ++ran;
        if (State::get().failures_in_current) {
            ++failed;
            failed_names.push_back(full);
            std::printf("[  FAILED  ] %s\n", full.c_str());
        } else {
            std::printf("[       OK ] %s\n", full.c_str());
        }
        std::fflush(stdout);
    }
    if (list_flag()) return 0;
    std::printf("[==========] %d tests ran.\n[  PASSED  ] %d tests.\n", ran, ran - failed);
    if (failed) {
        std::printf("[  FAILED  ] %d tests, listed below:\n", failed);
        for (auto& n : failed_names) std::printf("[  FAILED  ] %s\n", n.c_str());
    }
    std::fflush(stdout);
    return failed ? 1 : 0;
}

}  // namespace testing

#define RUN_ALL_TESTS() ::testing::RunAllTests()

#define GTEST_SHIM_CLASS_(suite, name) suite##_##name##_Test

#define GTEST_SHIM_TEST_(suite, name, parent)                                                           \
    class GTEST_SHIM_CLASS_(suite, name) : public parent {                                              \
    public:                                                                                             \
        void TestBody() override;                                                                       \
    };                                                                                                  \
    static ::testing::internal::Registrar gtest_shim_reg_##suite##_##name(                              \
        #suite, #name, [] { GTEST_SHIM_CLASS_(suite, name) t; t.Run(); });                              \
    void GTEST_SHIM_CLASS_(suite, name)::TestBody()

#define TEST(suite, name) GTEST_SHIM_TEST_(suite, name, ::testing::Test)
#define TEST_F(fixture, name) GTEST_SHIM_TEST_(fixture, name, fixture)

#define GTEST_SHIM_AMBIGUOUS_ELSE_BLOCKER_ switch (0) case 0: default:

#define GTEST_SHIM_NONFATAL_(result_expr)                                              \
    GTEST_SHIM_AMBIGUOUS_ELSE_BLOCKER_                                                  \
    if (const ::testing::internal::Result gtest_shim_r = (result_expr)) ;               \
    else ::testing::internal::Reporter(__FILE__, __LINE__, gtest_shim_r.text) = ::testing::Message()

#define GTEST_SHIM_FATAL_(result_expr)                                                 \
    GTEST_SHIM_AMBIGUOUS_ELSE_BLOCKER_                                                  \
    if (const ::testing::internal::Result gtest_shim_r = (result_expr)) ;               \
    else return ::testing::internal::Reporter(__FILE__, __LINE__, gtest_shim_r.text) = ::testing::Message()

#define EXPECT_EQ(a, b) GTEST_SHIM_NONFATAL_(::testing::internal::cmp_eq(#a, #b, a, b))
#define EXPECT_NE(a, b) GTEST_SHIM_NONFATAL_(::testing::internal::cmp_ne(#a, #b, a, b))
#define EXPECT_LT(a, b) GTEST_SHIM_NONFATAL_(::testing::internal::cmp_lt(#a, #b, a, b))
#define EXPECT_LE(a, b) GTEST_SHIM_NONFATAL_(::testing::internal::cmp_le(#a, #b, a, b))
#define EXPECT_GT(a, b) GTEST_SHIM_NONFATAL_(::testing::internal::cmp_gt(#a, #b, a, b))
#define EXPECT_GE(a, b) GTEST_SHIM_NONFATAL_(::testing::internal::cmp_ge(#a, #b, a, b))
#define EXPECT_TRUE(c) GTEST_SHIM_NONFATAL_(::testing::internal::check_bool(#c, static_cast<bool>(c), true))
#define EXPECT_FALSE(c) GTEST_SHIM_NONFATAL_(::testing::internal::check_bool(#c, static_cast<bool>(c), false))
#define EXPECT_FLOAT_EQ(a, b) GTEST_SHIM_NONFATAL_(::testing::internal::cmp_float_eq(#a, #b, a, b))
#define EXPECT_NEAR(a, b, tol) GTEST_SHIM_NONFATAL_(::testing::internal::cmp_near(#a, #b, #tol, a, b, tol))
#define EXPECT_STREQ(a, b) GTEST_SHIM_NONFATAL_(::testing::internal::cmp_streq(#a, #b, a, b))

#define ASSERT_EQ(a, b) GTEST_SHIM_FATAL_(::testing::internal::cmp_eq(#a, #b, a, b))
#define ASSERT_NE(a, b) GTEST_SHIM_FATAL_(::testing::internal::cmp_ne(#a, #b, a, b))
#define ASSERT_LT(a, b) GTEST_SHIM_FATAL_(::testing::internal::cmp_lt(#a, #b, a, b))
#define ASSERT_LE(a, b) GTEST_SHIM_FATAL_(::testing::internal::cmp_le(#a, #b, a, b))
#define ASSERT_GT(a, b) GTEST_SHIM_FATAL_(::testing::internal::cmp_gt(#a, #b, a, b))
#define ASSERT_GE(a, b) GTEST_SHIM_FATAL_(::testing::internal::cmp_ge(#a, #b, a, b))
#define ASSERT_TRUE(c) GTEST_SHIM_FATAL_(::testing::internal::check_bool(#c, static_cast<bool>(c), true))
#define ASSERT_FALSE(c) GTEST_SHIM_FATAL_(::testing::internal::check_bool(#c, static_cast<bool>(c), false))
#define ASSERT_FLOAT_EQ(a, b) GTEST_SHIM_FATAL_(::testing::internal::cmp_float_eq(#a, #b, a, b))
#define ASSERT_NEAR(a, b, tol) GTEST_SHIM_FATAL_(::testing::internal::cmp_near(#a, #b, #tol, a, b, tol))
#define ASSERT_STREQ(a, b) GTEST_SHIM_FATAL_(::testing::internal::cmp_streq(#a, #b, a, b))

#define GTEST_SHIM_NO_THROW_(stmt, on_fail)                                                                 \
    GTEST_SHIM_AMBIGUOUS_ELSE_BLOCKER_                                                                       \
    if (::testing::internal::Result gtest_shim_r = [&]() -> ::testing::internal::Result {                    \
            try { stmt; } catch (const std::exception& e) {                                                 \
                return {false, std::string("Expected: " #stmt " doesn't throw an exception.\n  Actual: it throws \"") + e.what() + "\"."}; \
            } catch (...) {                                                                                  \
                return {false, "Expected: " #stmt " doesn't throw an exception.\n  Actual: it throws."};    \
            }                                                                                                \
            return {true, ""};                                                                              \
        }()) ;                                                                                               \
    else on_fail ::testing::internal::Reporter(__FILE__, __LINE__, gtest_shim_r.text) = ::testing::Message()

#define EXPECT_NO_THROW(stmt) GTEST_SHIM_NO_THROW_(stmt, )
#define ASSERT_NO_THROW(stmt) GTEST_SHIM_NO_THROW_(stmt, return)
