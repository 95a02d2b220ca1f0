// Compile-and-link check of the C++ host mirror (include/eagen_msm.hpp) against libeagen_msm.so.  Runs only calls that
// need no GPU (digit counts, FftPrecomp constants, the no-device error path) unless a device is present.
#include <cstdio>
#include "eagen_msm.hpp"
// -1 in Montgomery form without field arithmetic on this side: table_entry_by_id(base = 1, id = 1) = ((0 + 1) * -1)
static eagen::Felt minus_one() { return eagen::negbase_utils::table_entry_by_id(EAGEN_CURVE_PALLAS, 1, 1); }
int main() {
    using namespace eagen;
    if (argument_witness_calc::num_digits(EAGEN_CURVE_PALLAS, 5) != 56) return 1;
    if (argument_witness_calc::num_digits(EAGEN_CURVE_PALLAS, 2) != 129) return 2;
    Felt w = regular_functions_utils::FftPrecomp::omega_pow(EAGEN_CURVE_PALLAS, 32);   // omega^(2^32) = 1 (Montgomery R)
    Felt one = regular_functions_utils::FftPrecomp::half_pow(EAGEN_CURVE_PALLAS, 0);
    if (w != one) return 3;
    if (negbase_utils::id_by_digit(0).has_value() || *negbase_utils::id_by_digit(3) != 2 || negbase_utils::digit_by_id(2) != 3) return 4;
    // circuit-facing helpers (host side, no device)
    if (config::circuit_sizes(1000, 5) != std::make_pair((size_t)503, (size_t)503)) return 7;
    {
        // the Pasta generator is (-1, 2): -1 is an x-coordinate, so to_curve_x returns it unchanged and y_from_x / slope succeed
        Felt m1 = minus_one();
        Felt x = config::to_curve_x(EAGEN_CURVE_PALLAS, m1);
        if (x != m1) return 10;
        Felt y = config::y_from_x(EAGEN_CURVE_PALLAS, x);
        Felt s = config::slope(EAGEN_CURVE_PALLAS, x, y);
        if (y == Felt{0, 0, 0, 0} || s == Felt{0, 0, 0, 0}) return 11;
    }
    try {
        Context ctx(EAGEN_CURVE_PALLAS, 0);
        std::vector<Felt> sc(3, one);   // Montgomery "1" of Fp is not the scalar field's 1, but any value < 2^127 works... use zeros
        for (auto& s : sc) s = Felt{0, 0, 0, 0};
        std::vector<JacobianPoint> pts(3, JacobianPoint{});  // identities
        auto wit = argument_witness_calc::compute_lhs_witness(ctx, sc, pts, 5);
        if (wit.functions.size() != 56) return 5;
        auto entries = negbase_utils::prepare_scalar_witness(ctx, sc, 5, 56, 8, EAGEN_PSW_INTENDED);
        if (entries.size() != 3 * 5 * 8 || entries[0].kind != negbase_utils::Entry::Scalar) return 8;
        auto arr = regular_functions_utils::compute_divisor_witness_naive(ctx, pts);
        if (!arr.pos.empty() || !arr.neg.empty()) return 9;
        std::printf("gpu path ok: %zu functions\n", wit.functions.size());
    } catch (const Error& e) {
        if (e.status != EAGEN_E_NO_DEVICE) { std::printf("unexpected: %s\n", e.what()); return 6; }
        std::printf("no device: %s (expected on a CPU box; no fallback)\n", e.what());
    }
    return 0;
}
