"""Import-only stub (build container): tools/tsdf.py keeps an open3d HashSet of active voxels for its marching-cubes
extension; the TSDF integrate / sample arithmetic pinned by oracle/make_golden_tsdf.py never reads it."""


class _HashSet:
    def __init__(self, *a, **k):
        pass

    def insert(self, *a, **k):
        pass


class _Tensor:
    shape = (0,)

    @staticmethod
    def from_dlpack(x):
        return _Tensor()


class _Core:
    int64 = None
    Tensor = _Tensor

    @staticmethod
    def HashSet(*a, **k):
        return _HashSet()

    @staticmethod
    def Device(*a, **k):
        return None


core = _Core()
