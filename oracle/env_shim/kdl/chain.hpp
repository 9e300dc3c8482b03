#pragma once
#include "kdl_shim.hpp"
