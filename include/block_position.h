// Drop-in for the reference header of the same name (class BlockPosition): see include/bbme/dropin.hpp.
#pragma once
#include "bbme/dropin.hpp"
