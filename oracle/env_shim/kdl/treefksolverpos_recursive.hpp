#pragma once
#include "frames_io.hpp"
