class ECE:  # lib/metrics/utils.py:16 imports the name; only used at :270 (out of scope)
    def __init__(self, *a, **k):
        raise NotImplementedError("netcal is not installed; stub only")
