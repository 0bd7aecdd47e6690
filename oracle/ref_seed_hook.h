/*
 * oracle/ref_seed_hook.h -- force-included (-include) in front of the reference's translation units when they are built
 * into oracle/_ref.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference seeds a function-local `static std::mt19937 gen(rd())` from `std::random_device`
 * (include/helper/dim1algebra.hpp:2004,2070,2110).  To make runs repeatable WITHOUT touching the reference's sources,
 * the name is redirected to a stand-in that returns the seed chosen by oracle/ref_driver.cpp.
 */
#ifndef NPO_REF_SEED_HOOK_H
#define NPO_REF_SEED_HOOK_H
#include <random>
extern unsigned npo_ref_seed_value;
namespace std {
struct npo_seeded_device {
	typedef unsigned result_type;
	unsigned operator()() { return npo_ref_seed_value; }
};
} // namespace std
#define random_device npo_seeded_device
#endif
