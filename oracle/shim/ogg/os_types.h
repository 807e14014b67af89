#include "ogg.h"
