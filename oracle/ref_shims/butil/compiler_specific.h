#pragma once
#ifndef ALLOW_UNUSED
#define ALLOW_UNUSED __attribute__((unused))
#endif
#ifndef WARN_UNUSED_RESULT
#define WARN_UNUSED_RESULT __attribute__((warn_unused_result))
#endif
#include "butil/macros.h"
