"""TEST INFRASTRUCTURE ONLY -- astropy is absent; only Fast.save()/load() touch `fits`."""


class _Fits:
    def __getattr__(self, name):
        raise NotImplementedError("astropy stub (oracle shim)")


fits = _Fits()
