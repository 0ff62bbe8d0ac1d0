"""Stand-in for benchmark/testsystems/configuration.py:17-56.  The reference picks an OpenMM
Platform ("Reference", "CPU", "CUDA", "OpenCL") and, for CUDA, sets CudaPrecision=mixed and
DeterministicForces=true (:42-45).  This engine has exactly one platform -- sm_100a CUDA behind
liblsi.so -- so the object only records what the caller asked for and maps it to the arithmetic
mode: "Reference" -> all-fp64, anything else -> mixed (fp64 state and accumulators, fp32 forces).
Force summation is deterministic in both modes (every atom gathers its own force)."""


class Platform:
    def __init__(self, name="CUDA"):
        self._name = name
        self.precision = "double" if name == "Reference" else "mixed"
        self.properties = {"CudaPrecision": self.precision, "DeterministicForces": "true"}

    def getName(self):
        return self._name

    def getPropertyDefaultValue(self, key):
        return self.properties[key]

    def setPropertyDefaultValue(self, key, value):
        self.properties[key] = value
        if key in ("CudaPrecision", "OpenCLPrecision"):
            self.precision = "double" if value == "double" else "mixed"

    def __repr__(self):
        return "<lsi Platform '%s' (%s)>" % (self._name, self.precision)


def configure_platform(platform_name="Reference", fallback_platform_name="CPU"):
    return Platform(platform_name)
