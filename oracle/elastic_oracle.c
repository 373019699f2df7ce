/* placeholder, filled in below */
int oracle_elastic_placeholder(void) { return 0; }
