"""Host-side entity record with the reference's field names (optimax_rogue/game/entities.py:11-48).

No modifier or item subclass exists anywhere in the reference, so the derived stats
(game/attribles.py:21-43) equal the base stats and are plain properties here."""
import dataclasses


@dataclasses.dataclass
class Entity:
    iden: int
    depth: int
    x: int
    y: int
    health: int
    base_max_health: int
    base_damage: int
    base_armor: int

    @property
    def max_health(self):
        return self.base_max_health

    @property
    def damage(self):
        return self.base_damage

    @property
    def armor(self):
        return self.base_armor
