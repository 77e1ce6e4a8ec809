"""Test-infrastructure stub: lets `import torchmultimodal` succeed offline (torchmultimodal/utils/file_io.py:7
imports iopath only to download checkpoints, which the oracle never does)."""


class HTTPURLHandler:
    pass


class PathManager:
    def register_handler(self, handler):
        return None

    def get_local_path(self, path, **kwargs):
        import os
        if os.path.exists(path):      # local files pass through (e.g. a BPE merges file of the golden generators)
            return path
        raise RuntimeError("network access is not available in the oracle environment")

    def open(self, path, mode="r", **kwargs):
        return open(self.get_local_path(path), mode, **kwargs)
