#include "plugin.h"
