from .layers import resolve_activation
from .model import BaseModel, XPaiNN, load_model, resolve_model
from .output import resolve_output

__all__ = ["resolve_activation", "resolve_output", "resolve_model", "load_model", "BaseModel", "XPaiNN"]
