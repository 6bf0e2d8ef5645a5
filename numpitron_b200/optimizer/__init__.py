# ruff: noqa: F401
from .adam import Adam, AdamParams
