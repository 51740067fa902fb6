// cvshim: see opencv2/core/core.hpp
#include "../core/core.hpp"
