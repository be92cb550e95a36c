#pragma once
#include <cstddef>
#ifndef DISALLOW_COPY_AND_ASSIGN
#define DISALLOW_COPY_AND_ASSIGN(T) T(const T&) = delete; void operator=(const T&) = delete
#endif
#ifndef arraysize
template <typename T, size_t N> char (&SrArraySizeHelper(T (&array)[N]))[N];
#define arraysize(array) (sizeof(SrArraySizeHelper(array)))
#endif
