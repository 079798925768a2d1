// intentionally empty: serialize() members of the reference are never instantiated
