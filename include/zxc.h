/* zxc.h -- umbrella header (reference include/zxc.h:9-15). */
#ifndef ZXC_H
#define ZXC_H
#include "zxc_buffer.h"
#include "zxc_constants.h"
#include "zxc_dict.h"
#include "zxc_error.h"
#include "zxc_opts.h"
#include "zxc_pstream.h"
#endif /* ZXC_H */
