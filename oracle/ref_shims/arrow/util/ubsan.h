#pragma once
#include <cstring>
#include <type_traits>
namespace arrow { namespace util {
template <typename T> inline T SafeLoadAs(const uint8_t* unaligned) { T r; std::memcpy(&r, unaligned, sizeof(T)); return r; }
template <typename T> inline T SafeLoad(const T* unaligned) { T r; std::memcpy(&r, unaligned, sizeof(T)); return r; }
template <typename T> inline void SafeStore(void* unaligned, T value) { std::memcpy(unaligned, &value, sizeof(T)); }
}}
