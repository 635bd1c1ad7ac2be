"""Thin alias: the oracle engine lives in oracle/engine.py (test infrastructure)."""
from oracle.engine import OracleEngine, build_oracle_model  # noqa: F401
