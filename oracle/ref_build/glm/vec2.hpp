// GLM stand-in (test infrastructure only; see glm.hpp).
#pragma once
#include "glm.hpp"
