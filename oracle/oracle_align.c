#include "necat_oracle.h"
