/* placeholder replaced below */
int pto_version(void) { return 0; }
