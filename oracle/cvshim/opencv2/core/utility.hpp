// TEST INFRASTRUCTURE ONLY - see ../core.hpp
#pragma once
#include "../core.hpp"
