"""Enumerations shared by the decode path (values as in python_ldpc_app/enums.py:4-25)."""
from enum import Enum


class Result(Enum):
    """Outcome of ``SPA_Decoder.decode`` (spa_decoder.py:241,253)."""
    OK = "eOk"
    INVALID_INPUT = "eInvalidInput"
    INVALID_PATH = "eInvalidPath"
    DATA_TRANSFER_NOT_OK = "eDataTransferNotOk"

    @classmethod
    def from_flag(cls, converged) -> "Result":
        return cls.OK if converged else cls.DATA_TRANSFER_NOT_OK


class InterleaverType(Enum):
    NONE = "eNone"
    REGULAR = "eRegular"
    RANDOM = "eRandom"
    SRANDOM = "eSRandom"


class LDPCDecoderType(Enum):
    BIT_FLIPPING = "eBitFlipping"
    SUM_PRODUCT = "eSumProduct"


class EncodingMethod(Enum):
    STANDARD = "standard"
    RICHARDSON_URBANKE = "richardson_urbanke"
