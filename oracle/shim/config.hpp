// Generated equivalent of the reference's config.hpp.cmake.in (DC_MEM_ALIGNMENT default = 32).
#pragma once
#define DC_MEM_ALIGNMENT 32
