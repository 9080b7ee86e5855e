"""TEST INFRASTRUCTURE ONLY -- skyfield is absent; fast/complete_orbit_simulation.py:5-7
calls load.timescale() at import time."""


class _Load:
    def timescale(self):
        return None

    def tle_file(self, *args, **kwargs):
        raise NotImplementedError("skyfield stub (oracle shim)")


load = _Load()
wgs84 = None
