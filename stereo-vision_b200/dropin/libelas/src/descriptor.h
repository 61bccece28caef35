// Drop-in stub for libelas/src/descriptor.h: the descriptor stage (descriptor.cpp:28-121) runs on the
// device inside libelas_b200.so (k_descriptor).  The file exists only because stereomapper.pro lists it
// in HEADERS (stereomapper.pro:84).
#ifndef __DESCRIPTOR_H__
#define __DESCRIPTOR_H__
#endif
