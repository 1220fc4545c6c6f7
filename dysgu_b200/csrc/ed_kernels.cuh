#pragma once
#include "engine.cuh"
