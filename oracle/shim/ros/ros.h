// TEST INFRASTRUCTURE — the ROS logging / assertion macros the reference's factor sources use, as no-ops (messages) and aborts (assertions).
#pragma once
#include <cstdio>
#include <cstdlib>
#define ROS_INFO(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_DEBUG_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_ERROR_STREAM(x) ((void)0)
#define ROS_BREAK() std::abort()
#define ROS_ASSERT(c) do { if (!(c)) { std::fprintf(stderr, "ROS_ASSERT failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); std::abort(); } } while (0)
#define ROS_ASSERT_MSG(c, ...) ROS_ASSERT(c)
