// gtest_main.cpp -- the main() gtest_main would provide (the reference links GTest::gtest_main,
// /root/reference/CMakeLists.txt:49), for the stand-in header next to this file.
#include <gtest/gtest.h>

int main(int argc, char** argv) {
    ::testing::InitGoogleTest(&argc, argv);
    return RUN_ALL_TESTS();
}
