// Case alias: the reference includes "sort.h" (SplitBVHBuilder.cpp:5) but ships "Sort.h".
#include "Sort.h"
