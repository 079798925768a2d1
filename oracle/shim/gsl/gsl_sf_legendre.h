// intentionally empty (included by src/utility.h:16, never used)
