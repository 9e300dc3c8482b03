// stand-in: RealVectorSpace.h includes xArm6.h without using it (the real header needs KDL, urdfdom, glog and an STL reader)
#pragma once
#include "AbstractRobot.h"
