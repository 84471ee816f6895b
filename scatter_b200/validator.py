"""Checks on the loading dictionary -- behaviour of the reference's `scatter/validator.py:34-57`: an unknown load type
raises, `ini_steps` defaults to 5 (the dict is updated in place, as the reference does)."""

LOAD_TYPES = frozenset({"pulse", "heaviside", "moving", "moving_at_plane", "rose"})
DEFAULTS = {"ini_steps": 5}


class ValidateLoad:
    @staticmethod
    def validate(loading: dict) -> None:
        assert "type" in loading
        kind = loading["type"]
        if kind not in LOAD_TYPES:
            raise Exception(f'Error: Load type {kind} not supported')
        for key, value in DEFAULTS.items():
            loading.setdefault(key, value)
