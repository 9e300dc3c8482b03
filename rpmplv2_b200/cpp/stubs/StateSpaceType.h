#pragma once
namespace base {
// the reference's enum (include/state_spaces/StateSpaceType.h:13-20) plus the new value INTEGRATION.md adds
enum class StateSpaceType { Abstract, RealVectorSpace, RealVectorSpaceFCL, SO2, SO3, RealVectorSpaceCUDA };
}
