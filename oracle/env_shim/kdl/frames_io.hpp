// oracle/env_shim/kdl — see kdl_shim.hpp (work-alike of the orocos-KDL slice the reference's robot classes use).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include "kdl_shim.hpp"
