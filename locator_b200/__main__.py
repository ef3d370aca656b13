import sys

from .locator import main

sys.exit(main())
