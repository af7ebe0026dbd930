from lightning import Fabric  # noqa: F401
