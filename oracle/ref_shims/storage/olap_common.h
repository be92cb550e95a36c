#pragma once
#include "base/types/int128.h"
namespace starrocks {
typedef unsigned __int128 uint128_t;
}
