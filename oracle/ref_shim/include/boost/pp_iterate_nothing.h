/* intentionally empty: see preprocessor.hpp stub */
