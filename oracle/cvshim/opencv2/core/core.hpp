// oracle-build stand-in (see cvshim_core.hpp)
#include "../cvshim_core.hpp"
