// internal helpers shared by the translation units of libfastace_b200.so
#ifndef FASTACE_INTERNAL_H
#define FASTACE_INTERNAL_H
#include <string>
namespace fastace {
void set_error(const std::string& msg);
}
#endif
