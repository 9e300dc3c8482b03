#pragma once
#include "yaml.h"
