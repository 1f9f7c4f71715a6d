#pragma once
#include "simt.cuh"
namespace trpx {
}
