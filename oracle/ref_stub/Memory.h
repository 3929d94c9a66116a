// oracle/ref_stub/Memory.h -- TEST INFRASTRUCTURE ONLY.
//
// The reference's SimOuterSPACE.cpp:5 includes ramulator's <Memory.h>, a
// third-party DRAM timing model that is NOT vendored in the reference tree
// (simulator/Makefile:2,8,14 only names -Iramulator/ and libramulator.a; no
// version is pinned).  ramulator only feeds the cycle-level model
// (SimOuterSPACE.cpp:608-719); the functional TaskProvider (:44-144) never
// touches it.  This header declares just enough of the ramulator surface for
// that translation unit to compile, so oracle/ref_harness.cpp can run the
// unmodified TaskProvider.  Nothing here computes anything.
#pragma once
#include <functional>
#include <string>
#include <vector>

namespace ramulator {

class Config {
public:
    explicit Config(const std::string &) {}
    std::string operator[](const std::string &) const { return std::string(); }
    int get_channels() const { return 1; }
    int get_ranks() const { return 1; }
    void set_core_num(int) {}
};

class HBM {
public:
    enum class Level { Channel };
    HBM(const std::string &, const std::string &) {}
    void set_channel_number(int) {}
    void set_rank_number(int) {}
};

template <class T>
class DRAM {
public:
    DRAM(T *, typename T::Level) {}
    int id = 0;
    void regStats(const std::string &) {}
};

template <class T>
class Controller {
public:
    Controller(Config &, DRAM<T> *) {}
};

class Request {
public:
    enum class Type { READ, WRITE };
    Request(long, Type, std::function<void(Request &)>) {}
};

template <class T, template <class> class C>
class Memory {
public:
    Memory(Config &, std::vector<C<T> *>) {}
    bool send(Request) { return true; }
    void tick() {}
};

}  // namespace ramulator
