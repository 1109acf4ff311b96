/*
 * sws_numa.c -- NUMA placement of the host side of the transfer path (plain C, Linux sysfs + raw syscalls,
 * no libnuma).
 *
 * On a multi-socket box every GPU hangs off one socket's PCIe root; page-locked frames that live on the
 * other socket cross the inter-socket link on every DMA, and with one process per GPU that link -- not
 * PCIe, not the GPUs -- caps the aggregate host-frame throughput.  The reference has no counterpart (its
 * frames never leave host memory); the nearest thing is the slice-thread placement it leaves to the OS
 * (libavutil/slicethread.c).
 *   ff_b200_numa_node_of_pci()   NUMA node of a PCI device ("0000:1b:00.0"), -1 if unknown / single node
 *   ff_b200_numa_prefer()        allocate the calling thread's next pages on that node (MPOL_PREFERRED)
 *   ff_b200_numa_restore()       back to the default policy
 *   ff_b200_numa_bind_thread()   run the calling thread on the CPUs of that node
 * SWS_B200_NUMA=0 in the environment switches all of it off (A/B runs).
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <ctype.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "sws_internal.h"

#define MPOL_DEFAULT   0
#define MPOL_PREFERRED 1

static int numa_enabled(void)
{
    const char *e = getenv("SWS_B200_NUMA");
    return !(e && e[0] == '0');
}

int ff_b200_numa_node_of_pci(const char *bus_id)
{
    char path[128], id[32];
    FILE *f;
    int node = -1;
    size_t n;
    if (!bus_id || !numa_enabled())
        return -1;
    n = strlen(bus_id);
    if (n >= sizeof(id))
        return -1;
    for (size_t i = 0; i <= n; i++)
        id[i] = (char)tolower((unsigned char)bus_id[i]);
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", id);
    f = fopen(path, "r");
    if (!f)
        return -1;
    if (fscanf(f, "%d", &node) != 1)
        node = -1;
    fclose(f);
    return node;
}

int ff_b200_numa_prefer(int node)
{
    unsigned long mask[16] = { 0 };
    if (node < 0 || node >= (int)(sizeof(mask) * 8) || !numa_enabled())
        return -1;
    mask[node / (8 * sizeof(long))] |= 1UL << (node % (8 * sizeof(long)));
    return (int)syscall(SYS_set_mempolicy, MPOL_PREFERRED, mask, sizeof(mask) * 8);
}

void ff_b200_numa_restore(void)
{
    if (numa_enabled())
        syscall(SYS_set_mempolicy, MPOL_DEFAULT, NULL, 0);
}

/* "0-31,64-95" -> cpu_set_t */
static int parse_cpulist(const char *s, cpu_set_t *set)
{
    int count = 0;
    CPU_ZERO(set);
    while (*s) {
        char *end;
        long a = strtol(s, &end, 10), b;
        if (end == s)
            break;
        b = a;
        if (*end == '-') {
            s = end + 1;
            b = strtol(s, &end, 10);
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) {
            CPU_SET((int)c, set);
            count++;
        }
        s = *end == ',' ? end + 1 : end;
        if (*end != ',')
            break;
    }
    return count;
}

int ff_b200_numa_bind_thread(int node)
{
    char path[96], buf[1024];
    cpu_set_t want, have, both;
    FILE *f;
    int n = 0;
    if (node < 0 || !numa_enabled())
        return -1;
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f)
        return -1;
    if (!fgets(buf, sizeof(buf), f)) {
        fclose(f);
        return -1;
    }
    fclose(f);
    if (parse_cpulist(buf, &want) <= 0)
        return -1;
    /* stay inside whatever set the launcher (taskset, cgroup) already confined us to */
    if (sched_getaffinity(0, sizeof(have), &have) != 0)
        return -1;
    CPU_AND(&both, &want, &have);
    for (int c = 0; c < CPU_SETSIZE; c++)
        n += CPU_ISSET(c, &both) ? 1 : 0;
    if (n == 0)
        return -1;
    return sched_setaffinity(0, sizeof(both), &both) == 0 ? n : -1;
}
