// Placeholder for EnTT (un-vendored dependency of the reference, absent here).
// TEST INFRASTRUCTURE ONLY: lets objects/Entities.hpp parse so that the
// reference's voxels/Chunks.cpp can be compiled unmodified into oracle/_ref.
// Nothing on the light/mesh path touches the registry; every member aborts.
#pragma once
#include <cstdint>
#include <cstdlib>
namespace entt {
    enum class entity : std::uint32_t {};
    template <class... T> struct exclude_t {};
    template <class... T> inline constexpr exclude_t<T...> exclude {};
    struct null_t {
        template <class T> operator T() const { return T {}; }
    };
    inline constexpr null_t null {};
    template <class... T> struct view_stub {
        struct iterator {
            entity operator*() const { return entity {}; }
            iterator& operator++() { return *this; }
            bool operator!=(const iterator&) const { return false; }
        };
        iterator begin() const { return {}; }
        iterator end() const { return {}; }
        template <class F> void each(F&&) const {}
    };
    struct registry {
        template <class... A> bool valid(A&&...) const { return false; }
        template <class T, class... A> T& get(A&&...) const { std::abort(); }
        template <class T, class... A> T* try_get(A&&...) const { return nullptr; }
        template <class T, class... A> T& emplace(A&&...) { std::abort(); }
        template <class... T, class... A> view_stub<T...> view(A&&...) const { return {}; }
        template <class... A> entity create(A&&...) { return entity {}; }
        template <class... A> void destroy(A&&...) {}
        template <class... A> void clear(A&&...) {}
    };
}
