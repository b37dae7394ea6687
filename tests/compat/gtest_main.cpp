// Stand-in for libgtest_main (the reference links GTestMain, CMakeLists.txt:167).  See gtest/gtest.h.
#include "gtest/gtest.h"

int main(int argc, char **argv) {
  ::testing::InitGoogleTest(&argc, argv);
  return RUN_ALL_TESTS();
}
