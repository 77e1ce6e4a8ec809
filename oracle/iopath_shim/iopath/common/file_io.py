"""Test-infrastructure stub: lets `import torchmultimodal` succeed offline (torchmultimodal/utils/file_io.py:7
imports iopath only to download checkpoints, which the oracle never does)."""


class HTTPURLHandler:
    pass


class PathManager:
    def register_handler(self, handler):
        return None

    def get_local_path(self, path, **kwargs):
        raise RuntimeError("network access is not available in the oracle environment")
