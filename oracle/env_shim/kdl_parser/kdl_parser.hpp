#pragma once
#include "../kdl/frames_io.hpp"
