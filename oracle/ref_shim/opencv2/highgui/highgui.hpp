// TEST INFRASTRUCTURE ONLY (oracle/): see core/core.hpp.
#pragma once
#include <opencv2/core/core.hpp>
