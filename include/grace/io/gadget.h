// grace/io/gadget.h -- Gadget-2 (type 1) snapshot loader with the call shape of the reference's
// test helper read_gadget (tests/helper/read_gadget.cuh:69-167), SURVEY.md 8f N1.
#pragma once
#include "grace/device_vector.h"

#include <stdexcept>
#include <string>

namespace grace {

// Gas positions + smoothing lengths -> d_pos (resized to the gas count), float4 {x, y, z, h}.
// Throws std::runtime_error for an unreadable file or one without gas particles
// (read_gadget.cuh:85-90).
template <typename Float4Vec>
GRACE_HOST void read_gadget(const std::string& fname, Float4Vec& d_pos)
{
    long long n_gas = 0;
    if (grace_b200_gadget_info(fname.c_str(), nullptr, nullptr, &n_gas) != GRACE_B200_OK)
        throw std::runtime_error(grace_b200_last_error());
    if (n_gas == 0) throw std::runtime_error("Gadget file " + fname + " has no gas particles!");
    d_pos.resize((size_t)n_gas);
    size_t got = 0;
    if (grace_b200_read_gadget_f4(detail::context(), fname.c_str(), (float*)detail::raw(d_pos.data()), d_pos.size(),
                                  &got, nullptr) != GRACE_B200_OK)
        throw std::runtime_error(grace_b200_last_error());
}

} // namespace grace
