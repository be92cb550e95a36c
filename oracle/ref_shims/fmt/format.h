#pragma once
#include <string>
#include <cstddef>
namespace fmt {
template <typename... A> inline size_t formatted_size(const char*, const A&...) { return 0; }
template <typename O, typename... A> inline O format_to(O o, const char*, const A&...) { return o; }
template <typename... A> inline std::string format(const char*, const A&...) { return std::string(); }
}
