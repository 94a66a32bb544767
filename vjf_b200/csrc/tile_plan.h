// Constants shared by the device and the host side of the tile pipeline.
#pragma once
#define TK_MAXNS 8  // most stages of the layer-1 weight ring (mbarrier pairs reserved for them)
