// Test infrastructure: track::Idx_t (tracker/core/idx_t.h) reduced to the identity number Outline.h stores in DebugInfo.
#pragma once
#include <commons.pc.h>
namespace track {
struct Idx_t {
    uint32_t _identity = cmn::infinity<uint32_t>();
    constexpr Idx_t() = default;
    explicit constexpr Idx_t(uint32_t id) : _identity(id) {}
    constexpr uint32_t get() const { return _identity; }
    constexpr bool valid() const { return _identity != cmn::infinity<uint32_t>(); }
    constexpr auto operator<=>(const Idx_t& o) const = default;
};
}
