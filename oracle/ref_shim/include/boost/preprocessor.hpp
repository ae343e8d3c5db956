/* Stub for the oracle build: Boost.Preprocessor is absent from this image.
 * The reference's gb*mv.c use Boost.PP *file iteration* only to stamp out
 * fixed-bandwidth specialisations (kl==ku in 0..15) that dispatch ahead of a
 * general-bandwidth routine generated from the very same gb*mv.def template.
 * Making the iteration include nothing and the REPEAT expand to nothing
 * leaves an empty switch, so every call reaches the general routine: same
 * arithmetic, same loop order, only without compile-time bandwidths. */
#ifndef BOOST_PREPROCESSOR_HPP_STUB
#define BOOST_PREPROCESSOR_HPP_STUB
#define BOOST_PP_ITERATE() <boost/pp_iterate_nothing.h>
#define BOOST_PP_REPEAT_FROM_TO(a, b, macro, data)
#define BOOST_PP_INC(x) x
#define BOOST_PP_CAT(a, b) a ## b
#endif
