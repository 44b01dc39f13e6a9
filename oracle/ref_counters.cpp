// Storage for the call counters used by the -DCT_COUNT build of oracle/_ref (test infrastructure).
#ifdef CT_COUNT
unsigned long long g_ctBoxTests = 0, g_ctTriTests = 0;
#endif
