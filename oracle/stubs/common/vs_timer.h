// Stand-in for vscommon <common/vs_timer.h>: tic/toc live in the vs_common.h stub.
#pragma once
#include <common/vs_common.h>
