"""zstdsharp_b200 -- B200-native Zstandard batch codec behind the ZstdSharp API (hot path only).

Importing the package loads ``_build/libzstdb200.so``; it fails loudly if the library has not been built.
``zstdsharp_b200.datagen`` (pure numpy) can be imported on its own without the native library.
"""
__all__ = ["Compressor", "Decompressor", "MultiCodec", "ZstdException", "ZSTD_ErrorCode", "ZSTD_cParameter"]


def __getattr__(name):
    if name in __all__ or name in ("api", "_native"):
        import importlib
        api = importlib.import_module(".api", __name__)
        if name == "api":
            return api
        if name == "_native":
            return importlib.import_module("._native", __name__)
        return getattr(api, name)
    raise AttributeError(name)
