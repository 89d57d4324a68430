#include "xdr.h"
