#pragma once
#include <iostream>
struct SrNullStream { template <class T> SrNullStream& operator<<(const T&) { return *this; } };
#define DCHECK(x) if (false) SrNullStream()
#define DCHECK_EQ(a,b) if (false) SrNullStream()
#define DCHECK_NE(a,b) if (false) SrNullStream()
#define DCHECK_LE(a,b) if (false) SrNullStream()
#define DCHECK_LT(a,b) if (false) SrNullStream()
#define DCHECK_GE(a,b) if (false) SrNullStream()
#define DCHECK_GT(a,b) if (false) SrNullStream()
#define CHECK(x) if (false) SrNullStream()
#define CHECK_EQ(a,b) if (false) SrNullStream()
#define CHECK_LE(a,b) if (false) SrNullStream()
#define CHECK_LT(a,b) if (false) SrNullStream()
#define CHECK_GE(a,b) if (false) SrNullStream()
#define CHECK_GT(a,b) if (false) SrNullStream()
#define CHECK_NE(a,b) if (false) SrNullStream()
#define LOG(x) if (false) SrNullStream()
#define VLOG(x) if (false) SrNullStream()
