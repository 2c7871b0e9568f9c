#include "tf_shim.h"  /* stand-in so the reference TU compiles without TensorFlow; see oracle/ref_shim/tf_shim.h */
