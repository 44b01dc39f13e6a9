/* Force-included (-include) into the generated copy of the reference's bvh.cpp.
 * Test infrastructure only.  With -DCT_COUNT the hooks count IntersectAABB /
 * IntersectTriangle calls (bvh.cpp:165 / :147); otherwise they expand to nothing. */
#ifndef CT_REF_HOOKS_H
#define CT_REF_HOOKS_H
#ifdef CT_COUNT
extern unsigned long long g_ctBoxTests, g_ctTriTests;
#define CT_HOOK_BOX __atomic_fetch_add(&g_ctBoxTests, 1ull, __ATOMIC_RELAXED);
#define CT_HOOK_TRI __atomic_fetch_add(&g_ctTriTests, 1ull, __ATOMIC_RELAXED);
#else
#define CT_HOOK_BOX
#define CT_HOOK_TRI
#endif
#endif
