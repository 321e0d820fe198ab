"""Import stand-in for GPUtil (base_options.py:333 bestGPU): one idle GPU."""


class _G:
    memoryUtil = 0.0
    load = 0.0


def getGPUs():
    return [_G()]
