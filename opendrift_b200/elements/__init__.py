from .elements import LagrangianArray, DeviceElements  # noqa: F401
