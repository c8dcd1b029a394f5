"""reference path: upsnet/nms/nms.py -- gpu_nms_wrapper runs the device-resident sm_100a NMS."""
from upsnet_b200.operators import gpu_nms, gpu_nms_wrapper  # noqa: F401


def py_nms_wrapper(thresh):
    """nms/nms.py:26-29; same `IoU > thresh` semantics as the GPU kernel, so it maps onto it."""
    return gpu_nms_wrapper(thresh, 0)


cpu_nms_wrapper = py_nms_wrapper
