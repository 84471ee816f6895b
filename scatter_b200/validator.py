"""Load-dictionary validation -- mirror of the reference's `scatter/validator.py:4-57`."""
from typing import Dict

_KNOWN = ("pulse", "heaviside", "moving", "moving_at_plane", "rose")


class ValidateLoad:
    @staticmethod
    def validate(loading: Dict):
        """Checks the load type and fills in defaults (`ini_steps` = 5, validator.py:57)."""
        assert "type" in loading
        if loading["type"] not in _KNOWN:
            raise Exception(f'Error: Load type {loading["type"]} not supported')
        loading.setdefault("ini_steps", 5)
