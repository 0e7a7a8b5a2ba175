// oracle/shim/device_launch_parameters.h -- empty on the host (see cuda_runtime.h here).
#pragma once
