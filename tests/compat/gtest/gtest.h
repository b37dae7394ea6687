// Harness compatibility pack: the subset of googletest the reference's test/test.cpp uses
// (TEST, EXPECT_EQ / ASSERT_EQ / EXPECT_TRUE / ASSERT_TRUE and friends, RUN_ALL_TESTS), so that the file compiles
// UNCHANGED where googletest is not installed.  Output mimics gtest's summary lines.  tests/compat/gtest_main.cpp
// stands in for libgtest_main.
#ifndef CSB_COMPAT_GTEST_H
#define CSB_COMPAT_GTEST_H

#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing {

struct TestInfo {
  const char *suite, *name;
  void (*fn)();
};
inline std::vector<TestInfo> &registry() {
  static std::vector<TestInfo> r;
  return r;
}
inline int &current_failures() {
  static int f = 0;
  return f;
}
struct Registrar {
  Registrar(const char *s, const char *n, void (*fn)()) { registry().push_back(TestInfo{s, n, fn}); }
};
inline void InitGoogleTest(int *, char **) {}

template <typename A, typename B>
bool report_cmp(bool ok, const char *op, const char *ea, const char *eb, const A &a, const B &b, const char *file, int line) {
  if (ok) return true;
  std::ostringstream os;
  os << file << ":" << line << ": Failure\nExpected: (" << ea << ") " << op << " (" << eb << "), actual: " << a << " vs " << b << "\n";
  std::cerr << os.str();
  current_failures()++;
  return false;
}
inline bool report_bool(bool ok, const char *expr, bool want, const char *file, int line) {
  if (ok) return true;
  std::cerr << file << ":" << line << ": Failure\nValue of: " << expr << "\n  Expected: " << (want ? "true" : "false") << "\n";
  current_failures()++;
  return false;
}

}  // namespace testing

#define TEST(suite, name)                                                                       \
  static void suite##_##name##_Test();                                                          \
  static ::testing::Registrar suite##_##name##_registrar(#suite, #name, &suite##_##name##_Test); \
  static void suite##_##name##_Test()

// the comparison is done on the values as written (mixed signedness like the reference's EXPECT_EQ(340, v.size()))
#define CSB_GTEST_CMP_(a, b, op, fatal)                                                                         \
  do {                                                                                                          \
    const auto &csb_a_ = (a);                                                                                   \
    const auto &csb_b_ = (b);                                                                                   \
    if (!::testing::report_cmp((long long)csb_a_ op(long long) csb_b_, #op, #a, #b, csb_a_, csb_b_, __FILE__, __LINE__) && fatal) \
      return;                                                                                                   \
  } while (0)
#define EXPECT_EQ(a, b) CSB_GTEST_CMP_(a, b, ==, false)
#define ASSERT_EQ(a, b) CSB_GTEST_CMP_(a, b, ==, true)
#define EXPECT_NE(a, b) CSB_GTEST_CMP_(a, b, !=, false)
#define ASSERT_NE(a, b) CSB_GTEST_CMP_(a, b, !=, true)
#define EXPECT_LT(a, b) CSB_GTEST_CMP_(a, b, <, false)
#define EXPECT_LE(a, b) CSB_GTEST_CMP_(a, b, <=, false)
#define EXPECT_GT(a, b) CSB_GTEST_CMP_(a, b, >, false)
#define EXPECT_GE(a, b) CSB_GTEST_CMP_(a, b, >=, false)
#define EXPECT_TRUE(c) ::testing::report_bool(!!(c), #c, true, __FILE__, __LINE__)
#define EXPECT_FALSE(c) ::testing::report_bool(!(c), #c, false, __FILE__, __LINE__)
#define ASSERT_TRUE(c)                                                          \
  do {                                                                          \
    if (!::testing::report_bool(!!(c), #c, true, __FILE__, __LINE__)) return;   \
  } while (0)
#define ASSERT_FALSE(c)                                                         \
  do {                                                                          \
    if (!::testing::report_bool(!(c), #c, false, __FILE__, __LINE__)) return;   \
  } while (0)

inline int RUN_ALL_TESTS() {
  int failed = 0;
  std::vector<std::string> failed_names;
  printf("[==========] Running %zu tests.\n", ::testing::registry().size());
  for (const ::testing::TestInfo &t : ::testing::registry()) {
    printf("[ RUN      ] %s.%s\n", t.suite, t.name);
    fflush(stdout);
    ::testing::current_failures() = 0;
    t.fn();
    if (::testing::current_failures()) {
      failed++;
      failed_names.push_back(std::string(t.suite) + "." + t.name);
      printf("[  FAILED  ] %s.%s\n", t.suite, t.name);
    } else {
      printf("[       OK ] %s.%s\n", t.suite, t.name);
    }
    fflush(stdout);
  }
  printf("[==========] %zu tests ran.\n[  PASSED  ] %zu tests.\n", ::testing::registry().size(), ::testing::registry().size() - failed);
  if (failed) {
    printf("[  FAILED  ] %d tests, listed below:\n", failed);
    for (const std::string &n : failed_names) printf("[  FAILED  ] %s\n", n.c_str());
  }
  return failed ? 1 : 0;
}
#endif
