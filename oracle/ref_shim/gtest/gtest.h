/* Stand-in for <gtest/gtest.h> (googletest is not installed here): the reference's CPU sampling models
 * (cpp/tests/wholegraph_ops/graph_sampling_test_utils.cu) use these macros only as argument checks.
 * TEST INFRASTRUCTURE (oracle/_ref builds). */
#pragma once
#define EXPECT_EQ(a, b) ((void)0)
#define EXPECT_TRUE(a) ((void)0)
#define ASSERT_EQ(a, b) ((void)0)
#define ASSERT_TRUE(a) ((void)0)
#define FAIL() ((void)0)
namespace testing {
template <typename T>
class TestWithParam {
};
}  // namespace testing
